#!/bin/bash
# where run.sh's wall time goes: phases of the host build (RTC_TIMING) and the whole command, for three scenes
cd "$(dirname "$0")/../.."
for s in practice5_1 practice5_dragon_100k practice5_dragon_100k_glass; do
  for rep in 1 2; do
    t0=$(date +%s.%N)
    RTC_TIMING=1 ./run.sh scenes/$s.txt /tmp/cli_$s.ppm 2> /tmp/cli_$s.err
    t1=$(date +%s.%N)
    echo "$s run $rep: wall $(python -c "print('%.3f' % ($t1 - $t0))") s"
  done
  grep "rtc timing" /tmp/cli_$s.err | tr '\n' ';' | sed 's/\[rtc timing\]//g; s/  */ /g'; echo
done
