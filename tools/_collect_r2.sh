# round-2 artifact collection on the GPU box (one GPU): tests, per-kernel rooflines, bench (+reference arm), ncu launch list
# and full captures of the kernels this round touched, sanitizer.  Outputs under gpurun_out/r2f_*; condensed into profiles/
# here with tools/ncu_summary.py.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -n 4 gpurun_out/r2f_pytest_gpu.log
python -m pytest tests -m gpu -q -s -k "long_differential" 2>&1 | grep -E "differential|passed|failed" > gpurun_out/r2f_eq_differential.txt
python tools/kernel_roofline.py > gpurun_out/r2f_kernel_roofline.jsonl 2> gpurun_out/r2f_kr.err
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_bench_ref.json 2>> gpurun_out/r2f_bench.err
python tools/live_profile.py > gpurun_out/r2f_live_profile.jsonl 2>> gpurun_out/r2f_bench.err
python tools/live_profile.py --no-video >> gpurun_out/r2f_live_profile.jsonl 2>> gpurun_out/r2f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r2f_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:crossfade_flat -s 2 -c 1 -o gpurun_out/r2f_prof_crossfade python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r2f_ncu_crossfade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused -s 6 -c 2 -o gpurun_out/r2f_prof_fused python bench.py --workload audio --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r2f_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rgba_to_yuv -c 1 -o gpurun_out/r2f_prof_video python tools/kernel_roofline.py --only NOAUDIO --reps 1 > gpurun_out/r2f_ncu_video.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scale_tiled -c 8 -o gpurun_out/r2f_prof_scale python tools/kernel_roofline.py --only NOAUDIO --reps 1 > gpurun_out/r2f_ncu_scale.log 2>&1
for k in EqThree Envelope; do
  ncu --set full --clock-control none --import-source on -k regex:"eq_stream|envelope" -s 3 -c 1 -o gpurun_out/r2f_prof_big_$k python tools/kernel_roofline.py --only "$k" --reps 1 > gpurun_out/r2f_ncu_big_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 2 -c 1 -o gpurun_out/r2f_prof_resample python tools/kernel_roofline.py --only Resampler --reps 1 > gpurun_out/r2f_ncu_resample.log 2>&1
( echo "compute-sanitizer (memcheck, racecheck, synccheck) over the tests of the kernels written or rewritten in round 2";
  compute-sanitizer --tool memcheck python -m pytest tests/test_fused_voice.py tests/test_resampler.py -m gpu -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY";
  compute-sanitizer --tool memcheck python -m pytest tests/test_parity_audio.py tests/test_parity_video.py -m gpu -x -q -k "nvelope or eq_three_random or eq_three_golden or rgba or tiled_scaler or scal" 2>&1 | grep -E "passed|failed|ERROR SUMMARY";
  compute-sanitizer --tool racecheck python -m pytest tests/test_fused_voice.py tests/test_parity_audio.py -m gpu -x -q -k "config2_fused or odd_shapes or nvelope" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY";
  compute-sanitizer --tool synccheck python -m pytest tests/test_fused_voice.py tests/test_parity_audio.py -m gpu -x -q -k "config2_fused or nvelope" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" ) > gpurun_out/r2f_sanitizer.txt 2>&1
du -sh gpurun_out
