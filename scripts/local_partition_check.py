"""One process, one host thread, all devices: the single-process partition (gb25_exchange_connect_local) against the
single-GPU run, bit for bit, as the first thing a fresh process does with the library.  Usage:
    GB25_SYNC_TIMEOUT_S=20 python scripts/local_partition_check.py [ndev]
Dumps the Python stack and exits if the host thread blocks (faulthandler), so a deadlock costs a minute, not the box."""
import faulthandler
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    t0 = time.time()
    import gb25_b200  # noqa: F401
    from gb25_b200 import distributed as D
    print(f"import {time.time() - t0:.1f}s", flush=True)
    faulthandler.dump_traceback_later(int(os.environ.get("GB25_CHECK_WATCHDOG_S", "60")), exit=True)
    ok = True
    for gt in ("simple_lat_lon", "gaussian_islands"):
        t1 = time.time()
        r = D.local_partition_check(tuple(range(ndev)), gt, log=print)
        print(f"{gt}: {'bit-identical' if r else 'MISMATCH'} ({time.time() - t1:.1f}s)", flush=True)
        ok = ok and r
    faulthandler.cancel_dump_traceback_later()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
