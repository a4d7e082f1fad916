for f in batch-scheduler_b200/libbsched_*.so; do
  v=$(basename $f .so); v=${v#libbsched_}
  BS_LIB=$f timeout 120 python profiles/tools/fit_variants.py $v --cfg4-only >> gpurun_out/fv_sweep.jsonl 2>> gpurun_out/fv_sweep.err
done
cat gpurun_out/fv_sweep.jsonl
