python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/exp_tests.log
cat gpurun_out/exp_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e6_new.json 2> gpurun_out/e6_new.err
python - <<'PY'
import json
for n in ('new',):
    try:
        d=json.loads(open('gpurun_out/e6_%s.json'%n).read().strip().splitlines()[-1])
        f=d['roofline']['family_ms_per_step']
        print(n, round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:f[k] for k in ('gemm','gemm_hbm','layernorm','dwconv_gelu','window_attention')}, round(d['roofline']['frac'],3), round(d['roofline']['roofline_other']['frac'],3), d['stats'])
    except Exception as e: print(n,'ERR',e)
PY
