echo "=== quick bench" >> $LOG
timeout 600 python tools/quick_bench.py vgg300 32 >> $LOG 2>&1
timeout 600 python tools/quick_bench.py vgg300 64 >> $LOG 2>&1
tail -60 $LOG
