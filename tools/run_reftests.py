"""Run every tests/compat/_bin/reftest_N (the reference's tests/wrapper_test_N.cpp linked with the object our front-end
generated from tests/test_N.cpp) on the GPU box and write {N: "pass" | "fail: ..."} to the given JSON file."""
import glob, json, os, re, subprocess, sys, tempfile

here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
only = set(sys.argv[2:])
for exe in sorted(glob.glob(os.path.join(here, "tests", "compat", "_bin", "reftest_*"))):
    n = os.path.basename(exe)[len("reftest_"):]
    if only and n not in only:
        continue
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "build"))
        try:
            r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=60)
            txt = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout + r.stderr)
            # (a few wrappers only run the function and print nothing: exit status is their whole verdict)
            good = r.returncode == 0 and "failed" not in txt.lower() and (re.search(r"succeeded|passed|success", txt, re.I) or not txt.strip())
            out[n] = "pass" if good else f"fail: rc={r.returncode} " + txt.strip().replace("\n", " | ")[-200:]
        except subprocess.TimeoutExpired:
            out[n] = "fail: timeout"
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
p = sum(v == "pass" for v in out.values())
print(f"reference tests through the front-end: {p} pass / {len(out)} built")
for k, v in sorted(out.items(), key=lambda kv: (len(kv[0]), kv[0])):
    if v != "pass":
        print(k, v[:230])
