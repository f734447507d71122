python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/kernel_roofline.py --only Envelope 2>/dev/null | cut -c1-170
MXL_ENV_PER_THREAD=32 python tools/kernel_roofline.py --only Envelope 2>/dev/null | cut -c1-170
MXL_ENV_PER_THREAD=32 python -m pytest tests/test_parity_audio.py tests/test_parity_graph.py -m gpu -x -q -k "nvelope or all_audio" 2>&1 | tail -2
