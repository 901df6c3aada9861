DSPB_FORCE_G=16 python rtest3.py 2>&1 | grep "Gs/s"
