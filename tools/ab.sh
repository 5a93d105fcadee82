# development helper: A/B kernel variants built into gunship_rs_b200/variants/ on one GPU box
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --steps 8 --frames 2 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_$name.json'));print('$name',round(d['device_ms_per_step'],4),{k:v['ms'] for k,v in d['stages'].items()})"
}
V=$PWD/gunship_rs_b200/variants
for v in "$@"; do run $v GSC_LIB=$V/lib_$v.so; done
