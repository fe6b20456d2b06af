timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/r2_sortseg_c5_ncu.csv python profiles/exp_sortseg_c5_once.py > gpurun_out/r2_ncu27.log 2>&1
tail -2 gpurun_out/r2_ncu27.log; wc -l gpurun_out/r2_sortseg_c5_ncu.csv
