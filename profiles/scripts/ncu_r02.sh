# round-2 ncu captures (one GPU): launch list of the bench command, full captures of the headline kernel and of the
# thread-per-state-set Jacobian kernel.  Run each program plain first (must exit 0), then under ncu.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench_f64.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_ncu_launch.log 2>&1
python bench.py --grid 30 30 30 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --dtype f64 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:real_kernel -s 1 -c 1 -o gpurun_out/r02_real_f64 -f \
    python bench.py --grid 30 30 30 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --dtype f64 > gpurun_out/r02_ncu_real.log 2>&1
python bench_configs.py --only pulse_jac_64 --reps 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:pulsejac -s 1 -c 1 -o gpurun_out/r02_pulsejac_f64 -f \
    python bench_configs.py --only pulse_jac_64 --reps 1 > gpurun_out/r02_ncu_pj.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
