for d in 0 1 4 8 13; do NOTIMING=1 NOSAVE=1 B=9472 NDIR=2 IPN_GPF_DBG=$d timeout 120 python tests/dev/persist_time.py 2>&1 | grep "dbg=" | cut -c1-200; done
NOTIMING=1 NOSAVE=0 B=9472 NDIR=2 timeout 120 python tests/dev/persist_time.py 2>&1 | grep "dbg=" | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_tick_persist.py -x -q -m gpu 2>&1 | tail -1
