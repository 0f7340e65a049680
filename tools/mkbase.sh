#!/bin/bash
# tools/mkbase.sh [rev]: build tools/_base.so from a clean checkout of the git revision (default HEAD)
set -e
rev=${1:-HEAD}
rm -rf /tmp/basewt && mkdir -p /tmp/basewt
git archive $rev | tar -x -C /tmp/basewt
make -C /tmp/basewt lib > /tmp/basewt/build.log 2>&1
cp /tmp/basewt/stereomatching_b200/libstereo_b200.so tools/_base.so
