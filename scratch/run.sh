set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2y_smoke.log 2>&1; tail -2 gpurun_out/r2y_smoke.log
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -c 600 gpurun_out/r2y_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2y_ref.json 2> gpurun_out/r2y_ref.err
python - <<'PY'
import json
for f in ("gpurun_out/r2y_bench.json", "gpurun_out/r2y_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), json.dumps(d.get("configs", {}).get("sketch", {}))[:900])
    except Exception as e:
        print(f, "unreadable", e)
PY
