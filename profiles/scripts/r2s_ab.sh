# round-2 session 3: parity + timing of the real-valued kernels after a change: bash profiles/scripts/r2s_ab.sh [libs]
libs=${@:-libepgx}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2s_pytest.log; cat gpurun_out/r2s_pytest.log
for lib in $libs; do
  for dt in f64 f32; do
    EPGX_LIB=$PWD/epgpy_b200/$lib.so python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-e2e --dtype $dt > gpurun_out/r2s_${lib}_$dt.json 2>gpurun_out/r2s_${lib}_$dt.err; tail -c 300 gpurun_out/r2s_${lib}_$dt.err
  done
  EPGX_LIB=$PWD/epgpy_b200/$lib.so python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-e2e --max-nstate 32 > gpurun_out/r2s_${lib}_f64_n32.json 2>&1
  EPGX_LIB=$PWD/epgpy_b200/$lib.so python bench_configs.py --dtype f64 > gpurun_out/r2s_${lib}_cfg_f64.jsonl 2> gpurun_out/r2s_${lib}_cfg_f64.err
done
grep -H -o '"ms_per_step": [0-9.]*' gpurun_out/r2s_lib*.json
python - <<'PY'
import json,glob
for f in glob.glob('gpurun_out/r2s_lib*_cfg_f64.jsonl'):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split('/')[-1], d.get('config'), round(d.get('ms',0),3), round(d.get('fma_frac') or 0,3))
PY
