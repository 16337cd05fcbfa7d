# strong-scaling run on an 8-GPU box: bench at N = 1, 2, 4, 8, D2H scaling, NCCL tests, gather chunking at N = 8
set -x
python -m pytest tests/test_gpu_multirank.py tests/test_compat_reference.py -x -q -m gpu 2>&1 | tail -3
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    python profiles/scripts/d2h_scaling.py > gpurun_out/d2h_n$n.json 2>> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-extra > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n profiles/scripts/d2h_scaling.py > gpurun_out/d2h_n$n.json 2>> gpurun_out/scale_n$n.err
  fi
  tail -c 300 gpurun_out/scale_n$n.err
done
for ch in 2 4 16; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$ch bench.py --gpus 8 --steps 5 --warmup 3 --no-extra --no-e2e --gather-chunks $ch > gpurun_out/scale_n8_chunks$ch.json 2> gpurun_out/scale_n8_chunks$ch.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra --no-e2e --no-gather > gpurun_out/scale_n8_nogather.json 2> gpurun_out/scale_n8_nogather.err
