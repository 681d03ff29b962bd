python scripts/gpu_check.py 2>&1 | grep -E "bit-equal|rel err|NVE|force max" 
for t in "96,64,64,16,4" "96,64,128,16,4" "96,64,256,16,4" "96,64,128,16,2" "96,64,64,16,2" "96,64,128,16,1"; do
  CSS_TUNE="$t,768,448,1024,128,1" python scripts/perf_probe.py cfg5_torus_1Mfaces_N100k 2>&1 | tail -1
done
CSS_TUNE="96,64,128,16,4,768,448,1024,128,1" python scripts/perf_probe.py cfg4_icosphere_250kfaces_N25k 2>&1 | tail -1
