cd $GRAFT_REPO_ROOT
for c in 32768 131072 640000; do echo "chunk $c"; python bench.py --steps 4 --warmup 3 --no-cpu-baseline --ray-chunk $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"; done
