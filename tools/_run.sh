python -m pytest tests/test_parity_video.py -m gpu -x -q 2>&1 | tail -2
python tools/kernel_roofline.py --only NOAUDIO 2>/dev/null | tail -2 | cut -c1-170
