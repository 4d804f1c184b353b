python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "train or stft" 2>&1 | tail -8 > gpurun_out/exp_tests.log
cat gpurun_out/exp_tests.log
python tools/stft_bench.py > gpurun_out/stft_bench_new.json 2> gpurun_out/stft_bench_new.err
cat gpurun_out/stft_bench_new.json; tail -3 gpurun_out/stft_bench_new.err
