#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== index tests"; python -m pytest tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -2
python tools/perf_screen.py --steps 20 300 --out gpurun_out/r2d_ab.json --variants "stages6:" "stages5:AVL_MAX_STAGES=5"
python tools/perf_screen.py --steps 20 --variants "f16_noemit:AVL_DEBUG_FLAGS=16" "f4_nodrain:AVL_DEBUG_FLAGS=4"
AVL_DEBUG_FLAGS=64 python tools/perf_screen.py --child 20 2> gpurun_out/r2d_cta_full.txt > /dev/null; tail -150 gpurun_out/r2d_cta_full.txt | grep "avl clock" | tail -1
AVL_DEBUG_FLAGS=68 python tools/perf_screen.py --child 20 2> gpurun_out/r2d_cta_nodrain.txt > /dev/null
AVL_DEBUG_FLAGS=69 python tools/perf_screen.py --child 20 2> gpurun_out/r2d_cta_tmaonly.txt > /dev/null
AVL_DEBUG_FLAGS=70 python tools/perf_screen.py --child 20 2> gpurun_out/r2d_cta_mmaonly.txt > /dev/null
for f in full nodrain tmaonly mmaonly; do echo "== $f"; tail -149 gpurun_out/r2d_cta_$f.txt | grep "avl c" | python -c "
import sys
rows=[l.split() for l in sys.stdin if 'avl cta' in l]
ends=sorted(int(r[11]) for r in rows); durs=sorted(int(r[8]) for r in rows); mhz=sorted(float(r[13]) for r in rows)
print('n',len(rows),'end min/med/max',ends[0],ends[len(ends)//2],ends[-1],'dur min/med/max',durs[0],durs[len(durs)//2],durs[-1],'MHz',mhz[0],mhz[len(mhz)//2],mhz[-1])
"; done
