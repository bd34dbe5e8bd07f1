#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== index + dropin + p2p tests"; python -m pytest tests/test_index_gpu.py tests/test_dropin_gpu.py tests/test_handoff_p2p_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/perf_screen.py --steps 20 300 --out gpurun_out/r2i_ab.json --variants "dyn:"
python tools/perf_screen.py --steps 20 --variants "f16_noemit:AVL_DEBUG_FLAGS=16" "f4_nodrain:AVL_DEBUG_FLAGS=4"
AVL_DEBUG_FLAGS=64 python tools/perf_screen.py --child 20 2> gpurun_out/r2i_cta_full.txt > /dev/null
tail -149 gpurun_out/r2i_cta_full.txt | python -c "
import sys
rows=[l.split() for l in sys.stdin if 'avl cta' in l]
ends=sorted(int(r[12]) for r in rows); tiles=sorted(int(r[16]) for r in rows); mhz=sorted(float(r[14]) for r in rows)
print('n',len(rows),'end min/med/max',ends[0],ends[len(ends)//2],ends[-1],'tiles min/max',tiles[0],tiles[-1],'MHz',mhz[0],mhz[len(mhz)//2],mhz[-1])
"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-build --no-cpu --no-extra 2>/dev/null | cut -c1-330
