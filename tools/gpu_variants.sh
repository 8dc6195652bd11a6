# bench of tuning variants built with LCX_BUILD_TAG (lib_<tag>/): usage gpu_variants.sh tag1 tag2 ...   ("base" = the product)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = "base" ]; then unset LCX_B200_LIBDIR; else export LCX_B200_LIBDIR=$GRAFT_REPO_ROOT/libcloudphxx_b200/lib_$tag; fi
  python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/r02_var_$tag.json 2> gpurun_out/r02_var_$tag.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_var_$tag.json').read().strip().splitlines()[-1])
k=d['kernels']
print('$tag: ms/step %.3f  e2e %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']), {n: round(v['ms']/v['launches'],3) for n,v in list(k.items())[:6]})
P
done
