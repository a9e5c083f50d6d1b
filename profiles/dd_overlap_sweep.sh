# 4-GPU sweep of the tile-split (overlapped) layer schedule; results in profiles/experiments/README.md
cd $GRAFT_REPO_ROOT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 4 --steps 12 --warmup 4 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value']), round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()})"; }
GAMD_DD_OVERLAP=0 run 29601 "no-overlap            "
GAMD_DD_OVERLAP=1 run 29602 "overlap               "
GAMD_DD_OVERLAP=1 GAMD_DD_RESERVE_SMS=8 run 29603 "overlap reserve 8     "
# NCCL_P2P_USE_CUDA_MEMCPY=1 (copy-engine send/recv) HANGS with the torch-bundled NCCL 2.28.9 on this box: do not use
