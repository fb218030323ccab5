"""profiles/traffic.json from an ncu --set full capture of `tools/prof_case.py z2z 512 512 512` (runs without a GPU):
   python tools/make_traffic_json.py gpurun_out/prof.ncu-rep gpurun_out/prof_case_plain.log
dram__bytes_read.sum + dram__bytes_write.sum per launch, plus the fingerprint of the plan the capture was taken on
(sha1 of fftb200_describe, as bench.py computes it) so that bench.py refuses a stale capture."""
import csv, hashlib, io, json, os, subprocess, sys
rep, desc_log = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
desc = "\n".join(l.rstrip("\n") for l in open(desc_log) if l.startswith("tile ") or l.startswith("generic") or l.startswith("bluestein"))
out = {}
launches = [r for r in data if "fft" in r[ik]]
n = len([l for l in desc.strip().split("\n") if l])
for i, r in enumerate(launches[-n:]):            # the last exec's launches
    out["z2z_512_launch%d" % i] = int(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
out["_plan_fingerprint"] = hashlib.sha1(desc.strip().encode()).hexdigest()[:16]
out["_plan"] = desc.strip().split("\n")
out["_source"] = ("ncu --set full --clock-control none --import-source on, tools/prof_case.py z2z 512 512 512 (%s): "
                  "dram__bytes_read.sum + dram__bytes_write.sum per launch of the last exec" % os.path.basename(rep))
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
