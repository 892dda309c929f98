#!/bin/bash
# round 2, call AD: same-box A/B of the op sweep (quick): committed streaming kernel vs the separable instantiation
mkdir -p gpurun_out/r02ad
L=$PWD/stylegan-for-facerec_b200/csrc
for v in base new base2 new2; do
  lib=$L/libsg2_b200.so; [ ${v:0:4} = base ] && lib=$L/libsg2_b200_base.so
  SG2_B200_LIB=$lib timeout 600 python tools/opbench.py --quick --no-ref > gpurun_out/r02ad/op_$v.jsonl 2> gpurun_out/r02ad/op_$v.err
done
python - <<'PY'
import json
def load(f): return {(r['op'],r['dtype'],r['res'],r['C']):r for r in map(json.loads, open(f))}
A,B,A2,B2=[load(f'gpurun_out/r02ad/op_{v}.jsonl') for v in ('base','new','base2','new2')]
for k in A:
    if k[0].startswith('upfirdn2d') and k[2]>=64 and k[3]==64:
        print(k, A[k]['frac_of_hbm_peak'], A2[k]['frac_of_hbm_peak'], '->', B[k]['frac_of_hbm_peak'], B2[k]['frac_of_hbm_peak'])
PY
