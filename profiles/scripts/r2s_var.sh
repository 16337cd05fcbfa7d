# A/B of compile-time variants of the real kernel on ONE box: bash profiles/scripts/r2s_var.sh lib1 lib2 ...
for lib in "$@"; do
  EPGX_LIB=$PWD/epgpy_b200/$lib.so python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  for dt in f64 f32; do
    for rep in 1 2; do
      EPGX_LIB=$PWD/epgpy_b200/$lib.so python bench.py --steps 4 --warmup 3 --no-extra --no-cpu --no-e2e --dtype $dt > gpurun_out/r2v_${lib}_${dt}_$rep.json 2>gpurun_out/r2v_${lib}.err; tail -c 300 gpurun_out/r2v_${lib}.err
    done
  done
  EPGX_LIB=$PWD/epgpy_b200/$lib.so python bench.py --steps 4 --warmup 3 --no-extra --no-cpu --no-e2e --max-nstate 32 > gpurun_out/r2v_${lib}_f64n32_1.json 2>&1
done
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_*.json
