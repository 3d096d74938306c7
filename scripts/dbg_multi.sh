cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -v -i "warn" | tail -30
echo "exit: ${PIPESTATUS[0]}"
