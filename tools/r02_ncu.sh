#!/bin/bash
# round 2: ncu evidence.  Every profiled command first runs plain (exit code checked), then under ncu.
# launch list of the main line, then --set full of the dominant kernel of each BASELINE config.
O=gpurun_out/r02_ncu
mkdir -p $O
B="python bench.py --no-also --no-cpu --no-e2e --no-cusparse --steps 2 --warmup 3"
$B > $O/plain_poisson.json 2> $O/plain_poisson.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_poisson_auto.csv $B > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
$B --format csr > $O/plain_poisson_csr.json 2> $O/plain_poisson_csr.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_poisson_csr.csv $B --format csr > $O/ncu_launches_csr.log 2>&1
echo "launch list csr rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_dia_mul_dia -s 3 -c 1 -o $O/dia $B > $O/ncu_dia.log 2>&1; echo "dia rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_num_tiny -s 3 -c 1 -o $O/tiny $B --format csr > $O/ncu_tiny.log 2>&1; echo "tiny rc=$?"
$B --workload uniform --format ell > $O/plain_uniform.json 2> $O/plain_uniform.err &&
ncu --set full --clock-control none --import-source on -k regex:k_ell_mul_ell -s 3 -c 1 -o $O/ell $B --workload uniform --format ell > $O/ncu_ell.log 2>&1; echo "ell rc=$?"
$B --workload rmat --scale 20 > $O/plain_rmat20.json 2> $O/plain_rmat20.err &&
ncu --set full --clock-control none --import-source on -k regex:k_num_global2 -s 1 -c 1 -o $O/global2 $B --workload rmat --scale 20 > $O/ncu_global2.log 2>&1; echo "global2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_rmat20.csv $B --workload rmat --scale 20 > $O/ncu_launches_rmat20.log 2>&1; echo "launch list rmat20 rc=$?"
ls -la $O
