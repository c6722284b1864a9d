import json, sys
rows = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gemm_shapes.json"))
print("total ms", round(sum(r["ms"] for r in rows), 2), "launches", sum(r["count"] for r in rows))
for r in rows[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{r['M']:6d} {r['N']:6d} {r['K']:6d} a_t={r['a_t']} b_t={r['b_t']} n={r['count']:3d} ms={r['ms']:7.3f} us/call={1e3*r['ms']/r['count']:7.1f} TF={r['tflops']:7.1f}")
