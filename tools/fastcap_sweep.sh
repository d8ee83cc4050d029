for cap in 0 192 384 768; do
  echo "== MRP_HOST_FAST_CAP=$cap"
  MRP_HOST_FAST_CAP=$cap MRP_HOST_PROFILE=1 PROF_CBS=1 python tools/prof_ecbs.py 1000 2>&1 | grep -v "^\[mrp_host\] [0-9]* instances, [0-9]* lock-step iterations, [0-9]* nodes, [0-9]* replans, 0 exp" | grep "solved\|sum of costs\|mrp_host.*2[0-9][0-9] instances\|CBS 8x8" | head -12
done
