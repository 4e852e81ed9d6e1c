#!/bin/bash
mkdir -p gpurun_out
T=${1:-r3i}
timeout 400 python tools/check_lvis_variants.py > gpurun_out/${T}_lvis_variants.json 2> gpurun_out/${T}_lvis_variants.err; python - <<'P'
import json
d=json.load(open('gpurun_out/'+__import__('sys').argv[1]+'_lvis_variants.json')) if False else json.load(open('gpurun_out/r3i_lvis_variants.json'))
print({k:(v if not isinstance(v,dict) else v.get('full', v)) for k,v in d.items() if 'brdf' in k or k in ('v3','v2')})
P
tail -n 3 gpurun_out/${T}_lvis_variants.err
NF_BRDF_V3=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_zzz_gpu_reference_code.py -m gpu -q -x -p no:cacheprovider -k "learned or brdf or fused or stage_b or golden or relight" > gpurun_out/${T}_gputest_brdf3.log 2>&1; tail -n 3 gpurun_out/${T}_gputest_brdf3.log
