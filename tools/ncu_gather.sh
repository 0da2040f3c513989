#!/bin/bash
# DRAM sectors fetched per random access of the gather microbenchmark (what does one missing sector cost?)
M=dram__sectors_read.sum,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum
for cfg in "400 16" "12800 16" "12800 32" "12800 64" "12800 128"; do
  set -- $cfg
  ncu --metrics $M --clock-control none -k regex:gather2 -s 1 -c 1 --csv --log-file gpurun_out/g.csv \
    python -c "import sapling_b200 as S; S.gather_bench2($1<<20, 1<<26, $2, 1, 8, 1)" > /dev/null 2>&1
  echo "== footprint_MB=$1 gran=$2 accesses=$((1<<26))"
  grep -E '"(dram__|l1tex__|gpu__|lts__)' gpurun_out/g.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"'
done
