python tools/shape_shout.py 1 3 2>&1 | tail -1
TSGPU_TUNING=msm_slotted=0 python tools/shape_shout.py 1 3 2>&1 | tail -1
TSGPU_TABLE_WINDOW_BITS=19 python tools/shape_shout.py 1 3 2>&1 | tail -1
TSGPU_RED_SPAN=8 TSGPU_RED_MIN_SPANS=65536 python tools/shape_shout.py 1 3 2>&1 | tail -1
TSGPU_TUNING=msm_quad_tree=0 python tools/shape_shout.py 1 3 2>&1 | tail -1
