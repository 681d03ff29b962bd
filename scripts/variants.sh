# developer tool (GPU box): time the NVE step with each variant library built by scripts/build_variant.sh
for v in "" $VARIANTS; do
  if [ -z "$v" ]; then unset CSS_LIB_PATH; else export CSS_LIB_PATH=$PWD/curvedspacesim_b200/libvariant_$v.so; fi
  PROBE_STEPS=10 timeout 300 python scripts/perf_probe.py cfg5_torus_1Mfaces_N100k "variant=$v" 2>&1 | tail -1
done
