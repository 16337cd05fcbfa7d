set -x
python bench.py --grid 30 30 30 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --dtype f64 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:real_kernel -s 1 -c 1 -o gpurun_out/r2s_real_f64 -f \
    python bench.py --grid 30 30 30 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra --dtype f64 > gpurun_out/r2s_ncu_real.log 2>&1
ls -la gpurun_out/r2s_real_f64.ncu-rep
