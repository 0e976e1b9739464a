# launch list of the bench command + full-set capture of every kernel of the step (cfg2, N = 1)
CMD="python bench.py --steps 2 --warmup 1 --repeats 1 --no-e2e --no-tokens --no-extra-workloads --no-cpu --no-extractor"
$CMD > gpurun_out/r02_plain_short.json 2> gpurun_out/r02_plain_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
python profiles/tune_stage.py auto 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fill_ff|k_insert_dense|k_scan_dense|k_place_dense|k_pillar_walk|k_scatter_wide' -s 48 -c 6 -o gpurun_out/r02_full python profiles/tune_stage.py auto 16 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log; cat gpurun_out/plain.log
