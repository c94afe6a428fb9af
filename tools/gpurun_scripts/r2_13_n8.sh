set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_scaling.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version" | tail -12
