# strong-scaling run on an 8-GPU box: bench at N = 1, 2, 4, 8 (+ reference arm once), D2H scaling, NCCL tests
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
nvidia-smi topo -m > gpurun_out/topo8.txt
