import subprocess, time, sys, os
os.environ["LD_LIBRARY_PATH"] = "/root/repo/alis_b200"
os.chdir("/tmp")
def t(cmd):
    s = time.time(); subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL); return time.time() - s
for n in (2000, 8000):
    print("stepdump phc", n, "steps: %.2f s" % t(["/root/repo/tests/bin/stepdump_mine_cuda", "phc", str(n), "/tmp/x.bin"]))
subprocess.run(["cp", "/root/repo/tests/bin/phc_band_diagram_cuda", "/tmp/"])
for n in (2000, 8000):
    print("example phc", n, "steps: %.2f s" % t(["./phc_band_diagram_cuda", "1.2", str(n)]))
print("dipole_pml stepdump 300: %.2f s" % t(["/root/repo/tests/bin/stepdump_mine_cuda", "dipole_pml", "300", "/tmp/x.bin"]))
