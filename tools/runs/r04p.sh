mkdir -p gpurun_out
NCU="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
XFM_PROFILE_CONFIG=retrieval timeout 600 $NCU --log-file gpurun_out/r04p_launches_retrieval.csv python tools/profile_step.py > gpurun_out/r04p_ncu_retrieval.log 2>&1; echo "rc=$?"
XFM_PROFILE_CONFIG=pretrain timeout 600 $NCU --log-file gpurun_out/r04p_launches_pretrain.csv python tools/profile_step.py > gpurun_out/r04p_ncu_pretrain.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/r04p_launches_retrieval.csv | head -14
python tools/summarize_launches.py gpurun_out/r04p_launches_pretrain.csv | head -34
