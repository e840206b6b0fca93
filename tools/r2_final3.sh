#!/bin/bash
# last GPU check: class counts that are not a multiple of 4 (head padding), nothing else
O=gpurun_out/final3; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -q -s -k "class_count or three_classes or head_forward" > $O/unit.log 2>&1; echo "unit rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/unit.log
