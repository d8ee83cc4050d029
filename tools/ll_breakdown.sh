# one lane, launch-by-launch: kernel time vs time inside the replan call
for cap in 0 384; do
  echo "== MRP_HOST_FAST_CAP=$cap, one lane"
  MRP_HOST_LANES=1 MRP_HOST_FAST_CAP=$cap MRP_HOST_PROFILE=1 MRP_LL_PROFILE=1 python tools/prof_ecbs.py 1000 2> gpurun_out/ll_$cap.log | head -3
  python tools/ll_profile_summary.py gpurun_out/ll_$cap.log 2>&1 | cut -c1-420
done
