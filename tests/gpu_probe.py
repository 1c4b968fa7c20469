"""Box probe printed at the start of GPU sessions (not a test)."""
import os, subprocess, json
def sh(c):
    try: return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e: return "ERR %s" % e
print(json.dumps({"nproc": os.cpu_count(), "mem": sh("free -g | sed -n 2p"), "shm": sh("df -h /dev/shm | tail -1"),
  "gpu": sh("nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader"),
  "strip": sh("strip --version | head -1")}, indent=1))
