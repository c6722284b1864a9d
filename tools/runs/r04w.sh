mkdir -p gpurun_out
NCU="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
for c in nlvr vqa retrieval; do
XFM_PROFILE_CONFIG=$c timeout 600 $NCU --log-file gpurun_out/r04w_launches_$c.csv python tools/profile_step.py > gpurun_out/r04w_ncu_$c.log 2>&1; echo "$c rc=$?"
python tools/summarize_launches.py gpurun_out/r04w_launches_$c.csv 2>/dev/null | head -14
done
