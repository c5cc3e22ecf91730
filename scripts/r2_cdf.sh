python -m pytest tests -m gpu -x -q -k "cdf or odd_k or above_64 or config4" 2>&1 | tail -3
python bench.py --config 4 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2_cdf_new.json 2> gpurun_out/r2_cdf_new.err; tail -c 300 gpurun_out/r2_cdf_new.err
NBK_CDF_KEYS=64 python bench.py --config 4 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2_cdf_old.json 2> gpurun_out/r2_cdf_old.err
python - <<'PY'
import json
for f in ("new", "old"):
    d = json.loads(open(f"gpurun_out/r2_cdf_{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.4f G q/s, %.1f ms/step" % (d["value"] / 1e9, d["ms_per_step"]), d["parity_sample"])
PY
