set -x
python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-extra > gpurun_out/scale2_n$n.json 2> gpurun_out/scale2_n$n.err
  tail -c 300 gpurun_out/scale2_n$n.err
done
