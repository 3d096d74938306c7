# per-kernel breakdown of one BASELINE config: CFG=cfg4 bash scripts/gpu_cfg_breakdown.sh
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CFG=${CFG:-cfg4}
timeout 600 python bench.py --config $CFG --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err; tail -2 gpurun_out/bench_$CFG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$CFG.json').read().strip().splitlines()[-1])
print(d['config'].get('workload'), round(d['ms_per_step'],2), 'sparse', round(d.get('sparse',{}).get('ms_per_step',0),2))
tot=sum(v['launches']*v['avg_ms'] for v in d['kernels'].values())
print({k:(v['launches'], round(v['avg_ms'],3), round(100*v['launches']*v['avg_ms']/tot,1)) for k,v in d['kernels'].items()})
PY
