"""One-off source transformation used when programmatic dependent launch was introduced (kept for the record, idempotent):
   K<<<G, B, S, ST>>>(ARGS);  ->  MSX_CUDA(msx_launch(K, dim3(G), dim3(B), S, ST, ARGS));
   and pdl_entry(); as the first statement of every __global__ function of the converted files."""
import re, sys

def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{": depth += 1
        elif ch in ")]}": depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out

def convert(src):
    out, i = "", 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out += src[i:]; break
        # kernel expression: walk back over identifier + optional template args
        k = j
        if src[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                if src[k] == ">": depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0: break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] == "_"): k -= 1
        kern = src[k:j]
        e = src.find(">>>", j)
        cfg = split_top(src[j + 3:e])
        assert len(cfg) == 4, (kern, cfg)
        assert src[e + 3] == "(", src[e:e + 40]
        depth, m = 0, e + 3
        while True:
            if src[m] == "(": depth += 1
            elif src[m] == ")":
                depth -= 1
                if depth == 0: break
            m += 1
        args = src[e + 4:m]
        pass
        out += src[i:k] + "MSX_CUDA(msx_launch(%s, dim3(%s), dim3(%s), %s, %s, %s))" % (kern, cfg[0], cfg[1], cfg[2], cfg[3], args)
        i = m + 1
    return out

def add_entry(src):
    out, i = "", 0
    for mt in re.finditer(r"__global__", src):
        pass
    pos = 0
    while True:
        g = src.find("__global__", pos)
        if g < 0: break
        p = src.find("(", g)
        # skip __launch_bounds__(...) groups: the parameter list is the last (...) before '{' or ';'
        while True:
            depth, m = 0, p
            while True:
                if src[m] == "(": depth += 1
                elif src[m] == ")":
                    depth -= 1
                    if depth == 0: break
                m += 1
            nxt = m + 1
            while src[nxt] in " \n\t\\": nxt += 1
            if src[nxt] in "{;": break
            p = src.find("(", m)
        if src[nxt] == "{" and not src[nxt + 1:].lstrip(" \n\\").startswith("pdl_entry();"):
            src = src[:nxt + 1] + "\n  pdl_entry();" + src[nxt + 1:]
        pos = nxt + 1
    return src

for path in sys.argv[1:]:
    s = open(path).read()
    s2 = add_entry(convert(s))
    if s2 != s:
        open(path, "w").write(s2)
        print("converted", path)
