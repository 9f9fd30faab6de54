#!/bin/bash
# CLI timing variants on the configs[1] workload (run after tests/test_large_cli_gpu.py has left the files in /dev/shm)
D=/dev/shm/kuq_bench_r666000000_g2000_k31m15
EXE=krakenuniq_b200/bin/classify
ARGS="-d $D/database.kdb -i $D/database.idx -a $D/taxDB -M -r /tmp/speed.report"
FQ=$D/sample_8000000.fq
run() { echo "== $1"; shift; env KUQ_SPARSE_SLOTS=1073741824 KUQ_TIMING=1 "$@" 2>&1 | grep -a "timing\] pipeline\|timing\] fill\|sequences (" ; }
run "default -t 64 -o tmpfs"            $EXE $ARGS -t 64 -o $D/speed.kraken $FQ
run "-o /dev/null"                      $EXE $ARGS -t 64 -o /dev/null $FQ
run "OMP_WAIT_POLICY=passive"           env OMP_WAIT_POLICY=passive $EXE $ARGS -t 64 -o $D/speed.kraken $FQ
run "-t 32"                             $EXE $ARGS -t 32 -o $D/speed.kraken $FQ
run "-t 16"                             $EXE $ARGS -t 16 -o $D/speed.kraken $FQ
rm -f $D/speed.kraken
