# artifact collection on the GPU box: tests, per-kernel rooflines, sweep, bench (+reference arm), ncu launch list and captures
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
tools/pipe_rates.bin > gpurun_out/pipe_rates.jsonl
tools/pcie_peak.bin > gpurun_out/pcie_peak.jsonl
python tools/kernel_roofline.py > gpurun_out/kernel_roofline.jsonl 2> gpurun_out/kr.err
python tools/roofline_sweep.py > gpurun_out/mixer_sweep.jsonl 2> gpurun_out/sweep.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python bench.py --workload audio --ticks-per-step 1024 --steps 50 --no-cpu-baseline > gpurun_out/bench_audio1024.json 2>> gpurun_out/bench.err
python bench.py --workload audio --steps 50 > gpurun_out/bench_audio128.json 2>> gpurun_out/bench.err
python tools/live_profile.py > gpurun_out/live_profile.jsonl 2>> gpurun_out/bench.err
python tools/live_profile.py --no-video >> gpurun_out/live_profile.jsonl 2>> gpurun_out/bench.err
python bench.py --ticks-per-step 1 --steps 2000 --warmup 50 --no-cpu-baseline --e2e-steps 300 > gpurun_out/bench_live.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:crossfade_flat -s 2 -c 1 -o gpurun_out/prof_crossfade python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_crossfade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"eq_stream|oscillator|mixer_kernel|panner|meter" -s 10 -c 5 -o gpurun_out/prof_audio_stages python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --workload audio > gpurun_out/ncu_audio.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scale_tiled|compose_rgba" -s 9 -c 2 -o gpurun_out/prof_video python tools/kernel_roofline.py --only NOAUDIO --reps 2 > gpurun_out/ncu_video.log 2>&1
for k in "Oscillator(sine)" EqThree FmSine Envelope "Amplifier(+control)" Meter PcmSink; do
  n=$(echo $k | tr -dc 'A-Za-z')
  ncu --set full --clock-control none --import-source on -s 3 -c 1 -o gpurun_out/prof_big_$n python tools/kernel_roofline.py --only "$k" --reps 1 > gpurun_out/ncu_big_$n.log 2>&1
done
# memcheck / racecheck over the tests of this session's kernels and modules
( echo "compute-sanitizer (memcheck, racecheck): host-slice calls, StreamInput, Monitor / StreamOutput feeds, warp-per-slot meter, end-stop crossfade, scaler, compositor";
  compute-sanitizer --tool memcheck python -m pytest tests/test_host_slices.py tests/test_stream_input.py tests/test_monitor_feed.py -m gpu -x -q -k "not queue_capacity" 2>&1 | grep -E "passed|failed|ERROR SUMMARY";
  compute-sanitizer --tool memcheck python -m pytest tests/test_parity_audio.py tests/test_parity_video.py -m gpu -x -q -k "meter or end_stop or missing_layer or compose or tiled_scaler" 2>&1 | grep -E "passed|failed|ERROR SUMMARY";
  compute-sanitizer --tool racecheck python -m pytest tests/test_parity_audio.py tests/test_monitor_feed.py -m gpu -x -q -k "meter or video_jobs" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY" ) > gpurun_out/sanitizer2.txt 2>&1
du -sh gpurun_out
