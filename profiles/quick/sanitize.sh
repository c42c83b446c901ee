#!/bin/bash
# compute-sanitizer over the whole dense path without Python: the `dense` driver on the prepared 8-block file
# (make_expected.py), every case of cases.txt under memcheck, the first 3-D case of each estimator and the first projected case also under
# racecheck, initcheck and synccheck.  Usage on the GPU box: bash profiles/quick/sanitize.sh  (log: gpurun_out/sanitize.log)
cd "$(dirname "$0")/_bin" || exit 1
mkdir -p ../../../gpurun_out
log=../../../gpurun_out/sanitize.log
: > $log
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
run() {   # tool, case name, alg, driver tail
  local tool=$1 name=$2 alg=$3; shift 3
  echo "=== $tool $name" >> $log
  timeout 600 $CS --tool $tool --error-exitcode 99 --print-limit 20 ./dense del.out $name.san.raw $alg "$@" >> $log 2>&1
  local rc=$?
  if [ $rc -eq 0 ] && cmp -s $name.san.raw $name.exp; then echo "OK   $tool $name" | tee -a $log
  else echo "FAIL $tool $name (exit $rc)" | tee -a $log; fi
}
while read -r name alg tail; do run memcheck $name $alg $tail; done < cases.txt
for tool in racecheck initcheck synccheck; do
  while read -r name alg tail; do
    case $name in a_3d_tess|b_3d_cic|d_proj_narrow_z) run $tool $name $alg $tail;; esac
  done < cases.txt
done
grep -c "^OK" $log | sed 's/^/passed: /'
