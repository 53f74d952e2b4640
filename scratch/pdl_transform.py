"""One-off source transform: `k<<<grid, block, smem, stream>>>(args);` -> `lb_launch(k, grid, block, smem, stream, args);`
and `lb_pdl_enter();` as the first statement of every __global__ kernel (programmatic dependent launch)."""
import re, sys, pathlib

def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{": depth += 1
        if ch in ")]}": depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out

def transform_launches(src):
    out, i, n = "", 0, 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out += src[i:]; break
        # kernel expression: walk back over identifier and optional template args
        k = j
        if src[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                if src[k] == ">": depth += 1
                if src[k] == "<":
                    depth -= 1
                    if depth == 0: break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"): k -= 1
        name = src[k:j]
        e = src.find(">>>", j)
        cfg = split_top(src[j + 3:e])
        assert len(cfg) == 4, (name, cfg)
        assert src[e + 3] == "(", name
        depth, m = 0, e + 3
        while True:
            if src[m] == "(": depth += 1
            if src[m] == ")":
                depth -= 1
                if depth == 0: break
            m += 1
        args = src[e + 4:m]
        out += src[i:k] + "lb_launch(" + name + ", " + ", ".join(cfg) + (", " + args if args.strip() else "") + ")"
        i = m + 1
        n += 1
    return out, n

def insert_enter(src):
    out, i, n = "", 0, 0
    for mt in re.finditer(r"__global__", src):
        pass
    pos = 0
    while True:
        j = src.find("__global__", pos)
        if j < 0: break
        p = src.find("(", j)
        # skip __launch_bounds__(...) groups: the parameter list is the last (...) before '{'
        while True:
            depth, m = 0, p
            while True:
                if src[m] == "(": depth += 1
                if src[m] == ")":
                    depth -= 1
                    if depth == 0: break
                m += 1
            rest = src[m + 1:]
            stripped = rest.lstrip()
            if stripped.startswith("{"):
                brace = m + 1 + (len(rest) - len(stripped))
                break
            p = src.find("(", m + 1)
        src = src[:brace + 1] + "\n  lb_pdl_enter();" + src[brace + 1:]
        pos = brace
        n += 1
    return src, n

for f in sys.argv[1:]:
    path = pathlib.Path(f)
    s = path.read_text()
    s, a = transform_launches(s)
    s, b = insert_enter(s)
    path.write_text(s)
    print(f, "launches", a, "kernels", b)
