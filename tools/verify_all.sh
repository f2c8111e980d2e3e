#!/usr/bin/env bash
# Everything the round-end driver checks, in one go on a B200 box (about 4 minutes):
#   build -> CPU suite -> GPU parity suite -> smoke -> both bench arms (JSON lines into gpurun_out/)
# usage: bash tools/verify_all.sh [tag]
set -euo pipefail
cd "$(dirname "$0")/.."
tag="${1:-verify}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests -x -q -m "not gpu"
python -m pytest tests -x -q -m gpu
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > "gpurun_out/${tag}_reference_1gpu.json"
python bench.py --steps 5 --warmup 3 | tail -1 > "gpurun_out/${tag}_ont16s_1gpu.json"
for p in kmercount cluster reassign snpcluster pileup; do
  python bench.py --path "$p" --steps 3 --warmup 3 | tail -1 > "gpurun_out/${tag}_${p}_1gpu.json"
done
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for name in ("reference", "ont16s", "kmercount", "cluster", "reassign", "snpcluster", "pileup"):
    d = json.loads(open(f"gpurun_out/{tag}_{name}_1gpu.json").read())
    e2e = (d.get("e2e") or {}).get("value")
    print(f"{name:10s} value {d['value']:.4g} {d['unit']}  e2e {e2e:.4g}  roofline.frac {(d.get('roofline') or {}).get('frac')}")
PY
