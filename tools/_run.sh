python tools/kernel_roofline.py --only NOAUDIO 2>/dev/null | tail -2 | cut -c1-170
