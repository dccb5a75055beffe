SHORT="python bench.py --steps 2 --warmup 1 --prewarm 1024 --envs 65536 --no-cpu --no-obs"
for G in 32 64 128; do
  SF_L2_FETCH=$G ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum,gpu__time_duration.sum --clock-control none -k regex:sf_step_kernel -s 1025 -c 1 --csv --log-file gpurun_out/gran_$G.csv $SHORT > /dev/null 2>&1
  echo "gran $G"; grep -E "dram__bytes|lts__t_sectors|xbar2l1tex|gpu__time" gpurun_out/gran_$G.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
