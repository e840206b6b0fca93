#!/bin/bash
# build the .so locally (cross-compile), then run a command on the B200 box
set -e
cd /root/repo
python __graft_entry__.py | tail -1
exec /usr/local/graft/bin/gpurun "$@"
