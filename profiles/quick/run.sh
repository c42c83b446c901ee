#!/bin/bash
# A GPU check that needs seconds, not minutes (no Python): the `dense` driver on the prepared block file against the
# expected files of make_expected.py, byte for byte.  Usage on the GPU box: bash profiles/quick/run.sh
cd "$(dirname "$0")/_bin" || exit 1
mkdir -p ../../../gpurun_out
log=../../../gpurun_out/quick.log
: > $log
while read -r name alg tail; do
  if timeout 20 ./dense del.out $name.raw $alg $tail 2>> $log && cmp -s $name.raw $name.exp; then echo "OK   $name" | tee -a $log
  else echo "FAIL $name (differing bytes: $(cmp -l $name.raw $name.exp 2>/dev/null | wc -l))" | tee -a $log; fi
done < cases.txt
