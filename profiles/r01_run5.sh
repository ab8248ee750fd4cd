set -x
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -x -q 2>&1 | tail -15
for g in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --gather $g > gpurun_out/s6_bench2_$g.json 2> gpurun_out/s6_bench2_$g.err; cat gpurun_out/s6_bench2_$g.json; tail -5 gpurun_out/s6_bench2_$g.err
done
