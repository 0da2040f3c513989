#!/bin/bash
# DRAM sectors / L2 hit rate of the query kernel under different L2 eviction hints (and other env knobs).
# usage: bash tools/ncu_hints.sh <tag> "<ENV=VAL ...>" ...
TAG=$1; shift
M=dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum,lts__t_requests_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg ncu --metrics $M --clock-control none -k regex:kmer_query -s 4 -c 1 --csv --log-file gpurun_out/${TAG}_hints_$i.csv \
    python bench.py --steps 3 --warmup 3 --cpu-baseline none --e2e-steps 1 > /dev/null 2>&1
  echo "== $cfg" >> gpurun_out/${TAG}_hints.txt
  grep -E '"(dram__|lts__|l1tex__|gpu__)' gpurun_out/${TAG}_hints_$i.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' >> gpurun_out/${TAG}_hints.txt
done
cat gpurun_out/${TAG}_hints.txt
