"""Writes profiles/sass_<kernel>.txt: cuobjdump -sass of the headline kernels of the built objects (vision_b200/build/*.o), with an
instruction histogram in front (UTCHMMA / UTCBAR = tcgen05.mma / commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, FFMA2 = packed fp32 FMA,
LDGSTS = cp.async, ATOMS = shared atomics).  Runs without a GPU:  python tools/dump_sass.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "vision_b200", "build")
TARGETS = [  # (object, regex on the demangled function name, output tag)
    ("roi_ops.o", r"roi_align_line_kernel<7, 2, false>", "roi_align_line_kernel"),
    ("roi_ops.o", r"roi_align_line_kernel<7, 2, true>", "roi_align_line_kernel_multilevel"),
    ("roi_ops.o", r"roi_align_band_kernel<7, 2>", "roi_align_band_kernel"),
    ("roi_ops.o", r"roi_pool_plane_kernel<float, true>", "roi_pool_plane_kernel"),
    ("resize_stream.o", r"resize_aa_stream_kernel<__half, 10, 8>", "resize_aa_stream_kernel_f16"),
    ("deform_conv2d_tc.o", r"deform_conv2d_tc_kernel<__nv_bfloat16, 512, 4, 32>", "deform_conv2d_tc_kernel_bf16_bn512"),
    ("deform_conv2d_tc.o", r"deform_conv2d_tc3_kernel<128>", "deform_conv2d_tc3_kernel_fp32"),
    ("nms.o", r"bnms_mask_kernel<float4, 1, 8>", "bnms_mask_kernel"),
    ("roi_backward.o", r"roi_align_bwd_plane_atomic_kernel", "roi_align_bwd_plane_atomic_kernel"),
    ("roi_backward.o", r"roi_align_bwd_plane_fast_kernel", "roi_align_bwd_plane_fast_kernel"),
]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks, cur, name = {}, None, None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur = blocks.setdefault(name, [])
        elif cur is not None:
            cur.append(line)
    return blocks


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    cache = {}
    for obj, pat, tag in TARGETS:
        path = os.path.join(BUILD, obj)
        if path not in cache:
            cache[path] = functions(path)
        hit = None
        for mangled, lines in cache[path].items():
            dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            if re.search(re.escape(pat), dem):
                hit = (dem, lines)
                break
        if hit is None:
            print("not found:", pat)
            continue
        dem, lines = hit
        ins = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip() for l in lines if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        ops = collections.Counter()
        for l in ins:
            body = re.sub(r"^/\*[0-9a-f]+\*/\s*", "", l)
            body = re.sub(r"^@!?U?P\d+\s+", "", body)
            ops[body.split()[0].rstrip(";").split(".")[0]] += 1
        with open(os.path.join(ROOT, "profiles", f"sass_{tag}.txt"), "w") as f:
            f.write(f"# {dem}\n# object: vision_b200/build/{obj} (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)\n")
            f.write(f"# {len(ins)} SASS instructions; opcode histogram:\n")
            for op, n in ops.most_common():
                f.write(f"#   {op:12s} {n}\n")
            f.write("\n".join(ins) + "\n")
        key = {k: ops[k] for k in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "FFMA2", "HFMA2", "LDGSTS", "ATOMS", "LDS", "STS", "SHFL", "VOTE", "CREDUX", "REDG", "MULTIMEM") if ops[k]}
        print(f"{tag}: {len(ins)} instructions {key}")


if __name__ == "__main__":
    main()
