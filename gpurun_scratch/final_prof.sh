python -m pytest tests/test_gpu_kmer.py tests/test_gpu_seed.py -m gpu -x -q -k "ragged_deal or ragged_direct or flat_items" > gpurun_out/s4_t9.log 2>&1; tail -3 gpurun_out/s4_t9.log
mkdir -p /tmp/rep
for c in c2 c3 c4 c5 ragged; do
  if [ $c = ragged ]; then cmd="python profiles/run_ragged.py 10000000 100 150 31 1 2"; else cmd="python profiles/run_config.py $c 2"; fi
  ncu --set full --clock-control none --import-source on -k regex:'kmer_fast_kernel|seed_jit_kernel' -s 1 -c 1 -o /tmp/rep/$c $cmd > /tmp/rep/$c.log 2>&1
  { echo "# ncu --set full --clock-control none, one launch of the config's hot kernel at HEAD of round 2 ($cmd)"; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader | sed 's/^/# clocks right after the capture (sm, max sm, throttle reasons, W): /'; python profiles/ncu_key_metrics.py /tmp/rep/$c.ncu-rep; echo; echo "top stall sites:"; python profiles/ncu_top_stalls.py /tmp/rep/$c.ncu-rep 25; } > gpurun_out/r02f_ncu_$c.txt 2>&1
done
cp /tmp/rep/c2.ncu-rep gpurun_out/r02f_c2.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02f_bench_under_ncu.log 2>&1
ls -la gpurun_out/r02f_*
