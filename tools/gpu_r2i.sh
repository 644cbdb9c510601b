#!/bin/bash
mkdir -p gpurun_out
for m in 1 2; do
echo "=== SSDB_TC_MTU=$m"
SSDB_TC_MTU=$m timeout 300 python tools/layer_bench.py vgg300 64 split conv1_2 conv2 conv3 conv4 conv5 conv6 conv7 conv8 head0 head1 2>&1 | cut -c1-140
done
