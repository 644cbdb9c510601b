#!/bin/bash
echo "--- default"; python tools/wgrad_probe.py 2>&1 | grep -E "^B64" | cut -c1-120
echo "--- SSDB_TC_MTU=2 SSDB_WG_MTU=2"; SSDB_TC_MTU=2 SSDB_WG_MTU=2 python tools/wgrad_probe.py 2>&1 | grep -E "^B64" | cut -c1-120
echo "--- SSDB_WG_MTU=1"; SSDB_WG_MTU=1 python tools/wgrad_probe.py 2>&1 | grep -E "^B64" | cut -c1-120
echo "--- default again"; python tools/wgrad_probe.py 2>&1 | grep -E "^B64" | cut -c1-120
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
