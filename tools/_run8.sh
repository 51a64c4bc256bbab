mkdir -p gpurun_out /tmp/tp
export LD_LIBRARY_PATH=$PWD/baseline/_ref/bin:$LD_LIBRARY_PATH
for b in b200 b200o cuda; do mkdir -p /tmp/tp/$b; (cd /tmp/tp/$b && timeout 300 $OLDPWD/baseline/_ref/bin/test_procedural $b > log.txt 2>&1; echo "$b rc=$?"; grep -i "finished" log.txt | tail -1); cp /tmp/tp/$b/test_procedural.png gpurun_out/tp_$b.png; done
echo done
