python profiles/tune_stage.py auto 16 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/launches.csv python profiles/tune_stage.py auto 16 > gpurun_out/ncu.log 2>&1
cat gpurun_out/plain.log
