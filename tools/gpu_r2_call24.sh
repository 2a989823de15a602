#!/bin/bash
TAG=${1:-r2c24}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2e" > $OUT/pytest_c2e.log 2>&1; echo "pytest c2e rc=$?"; tail -3 $OUT/pytest_c2e.log
timeout 300 python tools/kbench.py --only bwd 2>&1 | grep -E "c2e" | tee $OUT/kbench_bwd.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_head.csv python - > $OUT/head.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
import cp360_b200
dev = torch.device('cuda', 0)
for w in (8, 7):
    c2e = cp360_b200.Cube2Equi(w)
    x = torch.randn(6 * 16, 1000, w, w, device=dev, requires_grad=True)
    gs = torch.randn(16, 2 * w, 4 * w, device=dev)
    for _ in range(2):
        x.grad = None
        c2e.to_equi_max(x).backward(gs)
torch.cuda.synchronize()
PY
python - <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/%s/launches_head.csv' % __import__('os').environ.get('TAG', 'r2c24'))) if len(r) > 5 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:70], r[-1])
PY
