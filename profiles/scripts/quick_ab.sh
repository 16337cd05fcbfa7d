# quick A/B of kernel builds on the GPU box: bash profiles/scripts/quick_ab.sh [lib names without .so]
libs=${@:-libepgx}
for lib in $libs; do
  for dt in f64 f32; do
    EPGX_LIB=$PWD/epgpy_b200/$lib.so python bench.py --steps 3 --warmup 2 --no-extra --no-cpu --no-e2e --parity-atoms 0 --dtype $dt > gpurun_out/ab_${lib}_$dt.json 2>gpurun_out/ab_${lib}_$dt.err; tail -c 200 gpurun_out/ab_${lib}_$dt.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/ab_${lib}_$dt.json').read().strip().splitlines()[-1])
print('$lib', '$dt', round(d['ms_per_step'],2), 'ms  frac', round(d['roofline']['frac'],4))
PY
  done
done
