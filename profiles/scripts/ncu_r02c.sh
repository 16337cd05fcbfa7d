# round-2 (third session) measurement set, one GPU: tests, default bench line, launch list of the bench command, full
# ncu capture of the headline kernel, secondary configurations.  Each program runs plain first, then under ncu.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02d_pytest.log
python bench.py > gpurun_out/r02d_bench_f64_default.json 2> gpurun_out/r02d_bench_f64_default.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02d_bench_plain.json 2> gpurun_out/r02d_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02d_bench_f64.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02d_ncu_launch.log 2>&1
python bench.py --grid 30 30 30 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --dtype f64 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:real_kernel -s 1 -c 1 -o gpurun_out/r02d_real_f64 -f \
    python bench.py --grid 30 30 30 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --dtype f64 > gpurun_out/r02d_ncu_real.log 2>&1
python bench_configs.py --dtype f64 > gpurun_out/r02d_cfg_f64.jsonl 2> gpurun_out/r02d_cfg_f64.err
python bench_configs.py --dtype f32 > gpurun_out/r02d_cfg_f32.jsonl 2> gpurun_out/r02d_cfg_f32.err
python bench.py --dtype f32 --no-cpu --no-extra > gpurun_out/r02d_bench_f32.json 2> gpurun_out/r02d_bench_f32.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02d_smoke.log 2>&1; tail -3 gpurun_out/r02d_smoke.log; cat gpurun_out/r02d_pytest.log; tail -c 600 gpurun_out/r02d_bench_f64_default.err
