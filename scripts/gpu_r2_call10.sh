#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; rm -f gpurun_out/parity_chain.jsonl
export NVSR_PARITY_REPORT=$GRAFT_REPO_ROOT/gpurun_out/parity_chain.jsonl
echo "=== A/B"; LIBS="ab_libs/lib_head.so neural-volume-super-resolution_b200/libnvsr_b200.so" bash scripts/gpu_ab_mlp.sh 2>&1 | grep -E "^==|NVSR_DBG" | tail -6
echo "=== tests"; timeout 1200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_parity_chain.py tests/test_gpu_train_tc.py tests/test_gpu_sr.py -m gpu -q --tb=short 2>&1 | tail -8
bash scripts/gpu_r2_bench_quick.sh
python - <<'PY'
import json, collections
agg=collections.defaultdict(lambda: collections.defaultdict(float))
for l in open('gpurun_out/parity_chain.jsonl'):
    d=json.loads(l)
    for k in ('coarse_unexplained_max','fine_tf_unexplained_max','coarse_sigma_maxdiff','fine_tf_sigma_maxdiff','coarse_logit_maxdiff','fine_tf_logit_maxdiff'):
        if k in d: agg[d['precision']][k]=max(agg[d['precision']][k], d[k])
for p,v in agg.items(): print(p, dict(v))
PY
