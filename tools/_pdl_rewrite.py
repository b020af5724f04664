"""One-off source transformation: PDL prologue in every kernel + launch_pdl() at every <<<>>> site."""
import re, sys

def match_fwd(s, i, op, cl):
    d = 0
    while i < len(s):
        if s[i] == op: d += 1
        elif s[i] == cl:
            d -= 1
            if d == 0: return i
        i += 1
    raise ValueError

def split_top(s):
    out, d, cur = [], 0, ""
    for ch in s:
        if ch in "([{": d += 1
        if ch in ")]}": d -= 1
        if ch == "," and d == 0:
            out.append(cur.strip()); cur = ""
        else: cur += ch
    out.append(cur.strip())
    return out

def rewrite(path, skip_kernels=()):
    s = open(path).read()
    # 1. kernels
    out, pos, nk = "", 0, 0
    for m in re.finditer(r"__global__", s):
        if m.start() < pos: continue
        # kernel name = identifier before the first '(' that follows (skip __launch_bounds__(...))
        i = m.end()
        while True:
            j = s.index("(", i)
            name = re.search(r"([A-Za-z_0-9]+)\s*$", s[i:j]).group(1)
            if name == "__launch_bounds__":
                i = match_fwd(s, j, "(", ")") + 1
                continue
            break
        close = match_fwd(s, j, "(", ")")
        brace = s.index("{", close)
        if s[close + 1:brace].strip() not in ("",):
            print("  odd kernel header", name); continue
        if name in skip_kernels or "pdl_sync();" in s[brace:brace + 200]:
            continue
        out += s[pos:brace + 1] + "\n  pdl_sync();"
        pos = brace + 1
        nk += 1
    s = out + s[pos:]
    # 2. launches
    out, pos, nl = "", 0, 0
    while True:
        k = s.find("<<<", pos)
        if k < 0: break
        # kernel expression before <<<
        b = k
        if s[b - 1] == ">":
            d, b = 0, b - 1
            while True:
                if s[b] == ">": d += 1
                elif s[b] == "<":
                    d -= 1
                    if d == 0: break
                b -= 1
        while b > 0 and (s[b - 1].isalnum() or s[b - 1] in "_:"): b -= 1
        kern = s[b:k]
        e = s.index(">>>", k)
        cfg = split_top(s[k + 3:e])
        while len(cfg) < 4: cfg.append("0")
        a0 = s.index("(", e)
        assert s[e + 3:a0].strip() == "", (path, s[e:a0 + 20])
        a1 = match_fwd(s, a0, "(", ")")
        args = s[a0 + 1:a1].strip()
        rep = f"SER_CUDA_CHECK(launch_pdl({kern}, dim3({cfg[0]}), dim3({cfg[1]}), {cfg[2]}, {cfg[3]}" + (", " + args if args else "") + "))"
        out += s[pos:b] + rep
        pos = a1 + 1
        nl += 1
    s = out + s[pos:]
    open(path, "w").write(s)
    print(path, "kernels", nk, "launches", nl)

if __name__ == "__main__":
    for p in sys.argv[1:]:
        rewrite(p)
