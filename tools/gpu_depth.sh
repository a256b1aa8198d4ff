#!/bin/bash
# streamed-gather pipeline depth of the cached CG kernel: expects variants/lib_d6.so, variants/lib_d8.so built with
#   make -C cmfrec_b200/csrc PREC=f32 EXTRA=-DCMF_RES_DEPTH=n   (after deleting build/obj/f32/sweep_cg_resident_m*.o)
qb() { timeout 300 python tools/quick_bench.py --shape $1 --k 64 --implicit $2 --iters 5 2>&1 | grep -E "RESULT|rror" | cut -c1-90; }
cp cmfrec_b200/lib/libcmfrec_b200_f32.so /tmp/orig.so
echo "== depth 4 (default)"; qb ml10m 0; qb lastfm 1
for d in 6 8; do cp variants/lib_d$d.so cmfrec_b200/lib/libcmfrec_b200_f32.so; echo "== depth $d"; qb ml10m 0; qb lastfm 1; done
cp /tmp/orig.so cmfrec_b200/lib/libcmfrec_b200_f32.so
