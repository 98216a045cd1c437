mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash scripts/gpu_run_tests.sh > gpurun_out/tests_all.log 2>&1
python bench.py > gpurun_out/bench_window.json 2> gpurun_out/bench_window.err
python bench.py --decode preset --no-cpu-baseline > gpurun_out/bench_preset.json 2> gpurun_out/bench_preset.err
python bench.py --workload stream --no-cpu-baseline > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
python bench.py --workload streams8 --no-cpu-baseline > gpurun_out/bench_streams8_1gpu.json 2> gpurun_out/bench_streams8_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1
bash scripts/gpu_profile.sh > gpurun_out/profile.log 2>&1
cat gpurun_out/summary.txt; tail -c 300 gpurun_out/bench_window.json; tail -c 300 gpurun_out/bench_streams8_1gpu.json; tail -5 gpurun_out/profile.log
