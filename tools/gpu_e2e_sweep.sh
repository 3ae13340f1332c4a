#!/bin/bash
# end-to-end call (pinned buffers, 10^6 problems) per chunk size / lanes / ramp
o=gpurun_out; tag=$1
run() {
  echo "== $*" >> $o/e2esweep_$tag.log
  env "$@" PIPLIB_B200_TIMING=0 timeout 200 python tools/e2e_timing.py 1000000 pinned 2>&1 | grep "^e2e" | tail -2 | cut -c1-60 >> $o/e2esweep_$tag.log
}
run A=0
run PIPLIB_B200_CHUNK=65536
run PIPLIB_B200_CHUNK=98304
run PIPLIB_B200_CHUNK=196608
run PIPLIB_B200_LANES=8
run PIPLIB_B200_LANES=8 PIPLIB_B200_CHUNK=98304
run PIPLIB_B200_LANES=4
run PIPLIB_B200_RAMP=2
run PIPLIB_B200_RAMP=4
run A=1
cat $o/e2esweep_$tag.log
