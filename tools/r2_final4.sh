#!/bin/bash
# last look at the final tree: the GPU suite only
O=gpurun_out/final4; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 170 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/all_gpu.log
