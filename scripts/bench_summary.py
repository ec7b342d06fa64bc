"""One line per bench block of a bench.py JSON line (read back from gpurun_out/).
   python scripts/bench_summary.py default.json [reference.json]      |      --brief "label" file.json"""
import json
import sys


def last_json(path):
    for ln in reversed(open(path).read().strip().splitlines()):
        ln = ln.strip()
        if ln.startswith("{"):
            return json.loads(ln)
    raise ValueError("no JSON line in " + path)


def show(name, b):
    m = b.get("measured", {})
    e = b.get("e2e") or {}
    p = b.get("parity_sample") or {}
    r = b.get("roofline") or {}
    c = b.get("cpu_baseline") or {}
    print(f"{name}: value {b['value']:.3e} e2e {e.get('value') or 0:.3e} kernel_ms {m.get('kernel_ms_per_launch', 0):.2f} "
          f"variant {m.get('kernel_variant')} frac {r.get('frac', 0):.3f} refills {m.get('corner_cache_refills_per_launch')} "
          f"deleted {m.get('deleted_per_launch')} | parity ok={p.get('ok')} n={p.get('n')} max_ulp={p.get('max_ulp')} "
          f"ei_mm={p.get('ei_mismatch')} state_mm={p.get('state_mismatch')} del {p.get('deleted_gpu')}/{p.get('deleted_oracle')} "
          f"| cpu {c.get('value') or 0:.3e} ({c.get('kind')}, {c.get('cores')} cores)")  # fmt: skip


def main():
    args = sys.argv[1:]
    if args and args[0] == "--brief":
        label, path = args[1], args[2]
        try:
            d = last_json(path)
            m = d.get("measured", {})
            e = d.get("e2e") or {}
            print(f"{label}: value {d['value']:.3e} kernel_ms {m.get('kernel_ms_per_launch', 0):.2f} e2e {e.get('value') or 0:.3e} "
                  f"ms/step {d.get('ms_per_step', 0):.2f}")  # fmt: skip
        except Exception as ex:  # noqa: BLE001
            print(label, "unreadable:", ex)
        return
    try:
        d = last_json(args[0])
        show(d["config"].get("workload", "main"), d)
        for k, v in d.get("extra", {}).items():
            show(k, v)
        if d.get("mode_d"):
            md = d["mode_d"]
            print("mode_d:", {k: md[k] for k in md if k not in ("config",)})
        if len(args) > 1:
            r = last_json(args[1])
            print("reference:", r.get("value"), r.get("cpu_baseline", r))
            print("same config:", d["config"] == r.get("config"))
    except Exception as ex:  # noqa: BLE001
        print("summary failed:", ex)


if __name__ == "__main__":
    main()
