#!/usr/bin/env python
"""apply_backend_patch.py — splice the B200 back end into the reference's KinematicChainMapping.

    python integration/apply_backend_patch.py <reference_root> <out_dir>

Reads   <reference_root>/src/sofa/RigidBodyDynamics/KinematicChainMapping.{h,inl}   (never modified)
Writes  <out_dir>/KinematicChainMapping.{h,inl}  (patched)  +  rbd_backend_glue.h, rbd_b200.h  (copied next to them)

The edits are described by NAME (function name + a token of its parameter list, brace-matched) and by short anchors,
and the inserted code comes from integration/KinematicChainMapping.backend.inl, so this repository holds none of the
reference's text: a unified diff would have to carry the removed reference lines.  Every edit must match exactly
once or the script fails (a changed reference is reported, not silently half-patched).

What changes (reference lines, src/sofa/RigidBodyDynamics/):
  KinematicChainMapping.h    + #include of the glue header; + members m_rbd / m_rbdRows / d_device / d_robots and the
                             stageBackend / fail / robots declarations (:119-128)
  KinematicChainMapping.inl  constructor (:48-49): two Data initialisers; init() (:62-95): stageBackend() at the end;
                             setModel (:304-311): no pinocchio::Data; the four protected workers (:338-396, :398-448,
                             :450-513, :515-639): bodies replaced by one rbd_*_host call each; + helper definitions
"""
from __future__ import annotations

import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REL = os.path.join("src", "sofa", "RigidBodyDynamics")


def load_sections(path: str) -> dict:
    sec, name, buf = {}, None, []
    with open(path) as f:
        for line in f:
            m = re.match(r"//@@ (\S+)\s*$", line)
            if m:
                if name:
                    sec[name] = "".join(buf).strip("\n") + "\n"
                name, buf = (m.group(1) if m.group(1) != "end" else None), []
            elif name:
                buf.append(line)
    return sec


def _skip_noncode(s: str, i: int) -> int:
    """If s[i:] starts a comment or a string / char literal, return the index just behind it, else i."""
    if s.startswith("//", i):
        j = s.find("\n", i)
        return len(s) if j < 0 else j
    if s.startswith("/*", i):
        j = s.find("*/", i + 2)
        return len(s) if j < 0 else j + 2
    if s[i] in "\"'":
        q, j = s[i], i + 1
        while j < len(s) and s[j] != q:
            j += 2 if s[j] == "\\" else 1
        return j + 1
    return i


def match_close(s: str, i: int, open_ch: str, close_ch: str) -> int:
    """s[i] == open_ch: index of the matching close_ch (comments and literals skipped)."""
    assert s[i] == open_ch
    depth, j = 0, i
    while j < len(s):
        k = _skip_noncode(s, j)
        if k != j:
            j = k
            continue
        if s[j] == open_ch:
            depth += 1
        elif s[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
        j += 1
    raise ValueError("unbalanced " + open_ch + close_ch)


def find_member_function(s: str, name: str, param_token: str | None):
    """(start of the parameter list '(', index of the body's '{', index of its '}') of the out-of-class definition of
    KinematicChainMapping<...>::name whose parameter list contains param_token (whitespace-insensitive)."""
    hits = []
    for m in re.finditer(r"KinematicChainMapping<\s*TIn\s*,\s*TInRoot\s*,\s*TOut\s*>::" + re.escape(name) + r"\s*\(", s):
        po = m.end() - 1
        pc = match_close(s, po, "(", ")")
        params = re.sub(r"\s+", "", s[po:pc + 1])
        if param_token is not None and re.sub(r"\s+", "", param_token) not in params:
            continue
        bo = pc + 1
        while s[bo] in " \t\r\n":
            bo += 1
        if s[bo] != "{":
            continue                                            # a declaration, not a definition
        hits.append((po, bo, match_close(s, bo, "{", "}")))
    if len(hits) != 1:
        raise SystemExit(f"expected exactly one definition of {name}({param_token or ''}), found {len(hits)}")
    return hits[0]


def replace_body(s: str, name: str, param_token, body: str) -> str:
    _, bo, bc = find_member_function(s, name, param_token)
    return s[:bo] + "{\n" + body + "  }" + s[bc + 1:]


def insert_after_line(s: str, pattern: str, text: str, within=None) -> str:
    lo, hi = within if within else (0, len(s))
    ms = [m for m in re.finditer(pattern, s) if lo <= m.start() < hi]
    if len(ms) != 1:
        raise SystemExit(f"anchor /{pattern}/ matched {len(ms)} times, expected 1")
    eol = s.find("\n", ms[0].end())
    return s[:eol + 1] + text + s[eol + 1:]


def insert_before_line(s: str, pattern: str, text: str) -> str:
    ms = list(re.finditer(pattern, s))
    if len(ms) != 1:
        raise SystemExit(f"anchor /{pattern}/ matched {len(ms)} times, expected 1")
    bol = s.rfind("\n", 0, ms[0].start()) + 1
    return s[:bol] + text + s[bol:]


def patch_header(h: str, sec: dict) -> str:
    if "rbd_backend_glue.h" in h:
        raise SystemExit("this tree is already patched (KinematicChainMapping.h includes rbd_backend_glue.h)")
    h = insert_after_line(h, r"#include\s*<pinocchio/multibody/model\.hpp>", sec["h.include"])
    h = insert_before_line(h, r"void\s+checkIndexFromRoot\s*\(\s*\)\s*;", sec["h.members"])
    return h


def patch_inl(s: str, sec: dict) -> str:
    # bodies first (they are found by name; later insertions do not disturb that)
    s = replace_body(s, "setModel", None, sec["inl.setModel.body"])
    s = replace_body(s, "apply", "VecCoord_t<Out>*out", sec["inl.apply.body"])
    s = replace_body(s, "applyJ", "VecDeriv_t<Out>*out", sec["inl.applyJ.body"])
    s = replace_body(s, "applyJT", "VecDeriv_t<In>*out", sec["inl.applyJT.body"])
    s = replace_body(s, "applyJT", "MatrixDeriv_t<In>*out", sec["inl.applyJT_matrix.body"])
    # constructor initialiser list: behind the (single-line) d_indexFromRoot initialiser
    s = insert_after_line(s, r":\s*d_indexFromRoot\s*\(\s*initData\s*\(", sec["inl.ctor_init"])
    # init(): stageBackend() after the base-class init
    _, bo, bc = find_member_function(s, "init", None)
    s = insert_after_line(s, r"Inherit::init\s*\(\s*\)\s*;", sec["inl.init_tail"], within=(bo, bc))
    # helper definitions in front of setModel's definition (its `template <...>` line)
    po, _, _ = find_member_function(s, "setModel", None)
    tpl = s.rfind("template", 0, po)
    bol = s.rfind("\n", 0, tpl) + 1
    s = s[:bol] + sec["inl.helpers"] + "\n" + s[bol:]
    return s


def main(argv):
    if len(argv) != 3:
        raise SystemExit(__doc__)
    ref, out = argv[1], argv[2]
    sec = load_sections(os.path.join(HERE, "KinematicChainMapping.backend.inl"))
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(ref, REL, "KinematicChainMapping.h")) as f:
        h = patch_header(f.read(), sec)
    with open(os.path.join(ref, REL, "KinematicChainMapping.inl")) as f:
        inl = patch_inl(f.read(), sec)
    with open(os.path.join(out, "KinematicChainMapping.h"), "w") as f:
        f.write(h)
    with open(os.path.join(out, "KinematicChainMapping.inl"), "w") as f:
        f.write(inl)
    shutil.copy(os.path.join(HERE, "rbd_backend_glue.h"), out)
    shutil.copy(os.path.join(HERE, "..", "include", "rbd_b200.h"), out)
    print(f"patched KinematicChainMapping.h / .inl written to {out}")


if __name__ == "__main__":
    main(sys.argv)
