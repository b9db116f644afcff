#!/bin/bash
# DNS: bucketed rank + fitness-sorted triangular k-NN: parity tests, timing at config 5, one ncu capture of the k-NN kernel
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k dns 2>&1 | tail -15
python tools/time_dns.py --check > gpurun_out/r1w_dns.json 2> gpurun_out/r1w_dns.err; cat gpurun_out/r1w_dns.json; tail -3 gpurun_out/r1w_dns.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1w_launches_dns.csv python tools/time_dns.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qdx_dns_knn -s 2 -c 1 -o gpurun_out/r1w_prof_dns -f python tools/time_dns.py > gpurun_out/r1w_ncu.log 2>&1
tail -2 gpurun_out/r1w_ncu.log
