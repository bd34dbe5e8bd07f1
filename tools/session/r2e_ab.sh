#!/bin/bash
# round 2, session e: per-CTA candidate buckets (no global atomics in the epilogue), device-side fallback, async call
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== index + dropin tests"; python -m pytest tests/test_index_gpu.py tests/test_dropin_gpu.py tests/test_handoff_p2p_gpu.py -x -q -m gpu 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/perf_screen.py --steps 20 300 --out gpurun_out/r2e_ab.json --variants "buckets:" "stages4:AVL_MAX_STAGES=4"
python tools/perf_screen.py --steps 20 --variants "f16_noemit:AVL_DEBUG_FLAGS=16" "f4_nodrain:AVL_DEBUG_FLAGS=4"
AVL_DEBUG_FLAGS=64 python tools/perf_screen.py --child 20 2> gpurun_out/r2e_cta_full.txt > /dev/null
for f in full; do echo "== $f"; tail -149 gpurun_out/r2e_cta_$f.txt | python -c "
import sys
rows=[l.split() for l in sys.stdin if 'avl cta' in l]
ends=sorted(int(r[12]) for r in rows); durs=sorted(int(r[9]) for r in rows); mhz=sorted(float(r[14]) for r in rows)
print('n',len(rows),'end min/med/max',ends[0],ends[len(ends)//2],ends[-1],'MHz',mhz[0],mhz[len(mhz)//2],mhz[-1])
"; done
