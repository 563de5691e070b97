set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02b_bench_n1.err
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02b_ncu_launches_step.csv python tests/diag/ncu_step.py > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
DV_NVTX=1 timeout 600 ncu --nvtx --print-nvtx-rename kernel --profile-from-start off --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/ncu_step.csv python tests/diag/ncu_step.py > gpurun_out/ncu_s.log 2>&1; echo "ncu traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tile -c 4 -o gpurun_out/r02b_conv_tile python tests/diag/ncu_big_layers.py > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_ -c 3 -o gpurun_out/r02b_bn python tests/diag/ncu_big_layers.py > gpurun_out/ncu_f2.log 2>&1; echo "ncu full bn rc=$?"
ls -la gpurun_out | tail -8
