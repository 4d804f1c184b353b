python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e2_base.json 2> gpurun_out/e2_base.err
WMK_DW_CPT=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e2_cpt2.json 2> gpurun_out/e2_cpt2.err
WMK_DW_CPT=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/exp_tests_cpt2.log
cat gpurun_out/exp_tests.log gpurun_out/exp_tests_cpt2.log
python - <<'PY'
import json
for n in ('base','cpt2'):
    try:
        d=json.loads(open('gpurun_out/e2_%s.json'%n).read().strip().splitlines()[-1])
        f=d['roofline']['family_ms_per_step']
        print(n, round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:f[k] for k in ('gemm','gemm_hbm','layernorm','dwconv_gelu','window_attention','small','layout')})
    except Exception as e: print(n,'ERR',e)
PY
