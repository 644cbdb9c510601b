#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_probe.py vgg300 64 2>&1 | tail -14
timeout 300 python tools/e2e_probe.py vgg512 32 2>&1 | tail -14
