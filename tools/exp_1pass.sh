for setting in "X=0" "MCQ_EPI_SKIP=1"; do
  echo "== $setting"; env $setting PROF_HW=16,64 timeout 120 python tools/prof_latency.py 2>&1 | grep "one_stream"
done
