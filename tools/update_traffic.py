"""profiles/traffic.json from `ncu --set full` reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels
bench.py quotes in `roofline.traffic`, stamped with the hash of the CUDA sources they were built from (bench.py's
kernel_source_hash: a later kernel change flags the entry stale instead of silently reusing it).
Usage: python tools/update_traffic.py loss.ncu-rep [wavelet.ncu-rep]     (no GPU needed: reads the reports with `ncu -i`)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def launches(rep):
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(r, name):
        return float(r[ix[name]].replace(",", "")) * scale[units[ix[name]]]

    return [(r[ix["Kernel Name"]], val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
             float(r[ix["gpu__time_duration.sum"]].replace(",", "")), units[ix["gpu__time_duration.sum"]]) for r in data]


def short_name(name):
    """'void wtpse::<unnamed>::db2_analysis_kernel<1, 1>(wtpse::<unnamed>::Db2FwdArgs)' -> 'db2_analysis_kernel<1, 1>'"""
    import re
    head = name.split("(")[0]
    m = re.search(r"([A-Za-z_0-9]+(?:<[^>]*>)?)\s*$", head)
    return m.group(1) if m else head


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    loss = launches(sys.argv[1])
    # bench.py's kernel names -> the template instance of the plain NCHW pair
    for key, pat in (("gram_tma_kernel", "gram_tma_kernel<0>"), ("apply_tma_kernel", "apply_tma_kernel(")):
        v = [b for name, b, _, _ in loss if pat in name]
        if v:
            t[key] = int(sum(v) / len(v))
    t["per_kernel"] = {}
    for name, b, _, _ in loss:
        t["per_kernel"].setdefault(short_name(name), []).append(int(b))
    t["per_kernel"] = {k: int(sum(v) / len(v)) for k, v in t["per_kernel"].items()}
    t["source"] = "ncu --set full --clock-control none of tools/loss_probe.py at 32x16x512x512 (%s): dram__bytes_read.sum + " \
                  "dram__bytes_write.sum per launch, mean over the captured launches" % os.path.basename(sys.argv[1])
    if len(sys.argv) > 2:
        wav = launches(sys.argv[2])
        per = {}
        for name, b, _, _ in wav:
            per.setdefault(short_name(name), []).append(b)
        t["wavelet_fused_step_32x2x512x512_db2_J4"] = int(sum(sum(v) / len(v) for v in per.values()))
        t["wavelet_per_kernel"] = {k: int(sum(v) / len(v)) for k, v in per.items()}
        t["wavelet_source"] = "ncu --set full of `bench.py --track wavelet` (%s): one launch of each kernel of the fused plan at " \
                              "32x2x512x512, db2, J=4" % os.path.basename(sys.argv[2])
    t["kernel_source_hash"] = bench.kernel_source_hash("whitening")
    if len(sys.argv) > 2:
        t["wavelet_source_hash"] = bench.kernel_source_hash("wavelet")
    t["hash_sources"] = {k: list(v) for k, v in bench.TRAFFIC_SOURCES.items()}
    json.dump(t, open(path, "w"), indent=1)
    print(json.dumps(t, indent=1))


if __name__ == "__main__":
    main()
