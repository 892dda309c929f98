#!/bin/bash
# same-box A/B of kernel variants: tools/ab.sh SIZE BATCH name1 "ENV1=.. ENV2=.." name2 "..." ...
SIZE=$1; BATCH=$2; shift 2
mkdir -p gpurun_out/ab
names=()
while [ $# -gt 0 ]; do
  name=$1; envs=$2; shift 2
  env $envs timeout 300 python bench.py --size $SIZE --batch $BATCH --no-cpu-baseline --steps 10 --warmup 3 \
      --profile-out gpurun_out/ab/${name}_$SIZE.json > gpurun_out/ab/${name}_$SIZE.log 2>&1 || echo "FAILED $name"
  names+=(gpurun_out/ab/${name}_$SIZE.json)
  grep -o '"value": [0-9.]*' gpurun_out/ab/${name}_$SIZE.log | head -1 | sed "s/^/$name $SIZE: /"
done
python tools/kcmp.py "${names[@]}" > gpurun_out/ab/table_$SIZE.txt
