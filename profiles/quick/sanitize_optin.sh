#!/bin/bash
# compute-sanitizer over the opt-in paths: TESSB200_FUSED=1 (k_cell_fused + k_cell_emit) and TESSB200_SEGMENTS=1 (per-point
# segments), memcheck + racecheck on the first 3-D cases.  Usage on the GPU box: bash profiles/quick/sanitize_optin.sh
cd "$(dirname "$0")/_bin" || exit 1
mkdir -p ../../../gpurun_out
log=../../../gpurun_out/sanitize_optin.log
: > $log
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
run() {   # env assignment, tool, case name, alg, driver tail
  local envs=$1 tool=$2 name=$3 alg=$4; shift 4
  echo "=== $envs $tool $name" >> $log
  env $envs timeout 900 $CS --tool $tool --error-exitcode 99 --print-limit 20 ./dense del.out $name.opt.raw $alg "$@" >> $log 2>&1
  local rc=$?
  if [ $rc -eq 0 ] && cmp -s $name.opt.raw $name.exp; then echo "OK   $envs $tool $name" | tee -a $log
  else echo "FAIL $envs $tool $name (exit $rc)" | tee -a $log; fi
}
while read -r name alg tail; do
  case $name in
    a_3d_tess) for tool in memcheck racecheck synccheck; do run TESSB200_FUSED=1 $tool $name $alg $tail; run TESSB200_SEGMENTS=1 $tool $name $alg $tail; done;;
    b_3d_cic|f_3d_narrow_xy) for tool in memcheck racecheck; do run TESSB200_SEGMENTS=1 $tool $name $alg $tail; done;;
  esac
done < cases.txt
grep -c "^OK" $log | sed 's/^/passed: /'
