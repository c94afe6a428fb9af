"""numpy interpreter of ``dmx_plan_dump`` output — test infrastructure.

Executes the *lowered* program (fused Pauli-transfer blocks, the device op table with its packed
coefficient tables, and the segment-kernel tables) exactly as the CUDA kernels are specified to, so the
host-side lowering in ``csrc/dmx_lower.cpp`` can be verified against the oracle on a machine without a GPU.
"""
from __future__ import annotations

import numpy as np


def _nums(s, typ=float):
    return np.array([typ(x) for x in s.split(",")] if s else [], dtype=typ)


def parse(dump: str) -> dict:
    plan = {"blocks": [], "passes": [], "phases": [], "sweeps": [], "sweeps3": [], "mus": [], "chains": [], "factors": [], "segs": [], "slots": {},
            "rots": [], "fsegs": [], "fclasses": [], "flabels": {}}
    for line in dump.splitlines():
        if not line.strip():
            continue
        toks = line.split(" ")
        head = toks[0]
        kv = {}
        pos = []
        for t in toks[1:]:
            if "=" in t:
                k, v = t.split("=", 1)
                kv[k] = v
            else:
                pos.append(t)
        if head == "plan":
            for k in ("n", "params", "path", "tile_k", "nseg", "cpb"):
                plan[k] = int(kv[k])
            plan["n_slots"] = int(kv["slots"])
        elif head == "block":
            b = {k: int(kv[k]) for k in ("kind", "nq", "q0", "q1", "cls")}
            for k in ("M", "A", "C", "S", "tab", "cond"):
                if k in kv:
                    b[k] = _nums(kv[k])
            for k in ("rot", "param", "slot"):
                if k in kv:
                    b[k] = int(kv[k])
            for k in ("scale", "offset"):
                if k in kv:
                    b[k] = float(kv[k])
            plan["blocks"].append(b)
        elif head == "pass":
            plan["passes"].append({"first_phase": int(kv["first_phase"]), "nphases": int(kv["nphases"]), "tile": _nums(kv["tile"], int),
                                   "pos_in": _nums(kv["pos_in"], int) if "pos_in" in kv else None,
                                   "pos_out": _nums(kv["pos_out"], int) if "pos_out" in kv else None})
        elif head == "phase":
            plan["phases"].append({k: int(kv[k]) for k in ("first_sweep", "nsweeps", "first_chain", "nchains")})
        elif head == "sweep":
            sw = {k: int(kv[k]) for k in ("p0", "p1", "first_mu", "nmu", "lperm", "lcoef", "sperm", "scoef")}
            sw["lmask"] = _nums(kv["lmask"], int)
            sw["smask"] = _nums(kv["smask"], int)
            plan["sweeps"].append(sw)
        elif head == "sweep3":
            sw = {k: int(kv[k]) for k in ("p0", "p1", "p2", "first_mu", "nmu", "lperm", "lcoef", "sperm", "scoef")}
            sw["lmask"] = _nums(kv["lmask"], int)
            sw["smask"] = _nums(kv["smask"], int)
            sw["wide"] = int(kv.get("wide", 0))
            if sw["wide"]:
                sw["lnvar"], sw["snvar"] = int(kv["lnvar"]), int(kv["snvar"])
                for key in ("ltmask", "stmask", "lvmask", "svmask"):
                    sw[key] = _nums(kv[key], int)
            plan["sweeps3"].append(sw)
        elif head == "mu":
            plan["mus"].append({k: int(kv[k]) for k in ("kind", "a0", "a1", "a2", "a3")})
        elif head == "chain":
            plan["chains"].append({"first": int(kv["first"]), "n": int(kv["n"])})
        elif head == "factor":
            plan["factors"].append({"kind": int(kv["kind"]), "coef": int(kv["coef"]), "rot": int(kv["rot"])})
        elif head == "" and "coefs" in kv:
            plan["coefs"] = _nums(kv["coefs"])
        elif head == "terms":
            plan["term_idx"] = _nums(kv["idx"], int)
            plan["term_coef"] = _nums(kv["coef"])
        elif head == "seg":
            plan["segs"].append({"rot": int(kv["rot"]), "src0": _nums(kv["src0"], int), "src1": _nums(kv["src1"], int),
                                 "active": _nums(kv["active"], int), "w0": _nums(kv["w0"]), "w1": _nums(kv["w1"])})
        elif head == "fseg":
            plan["fsegs"].append({"mask": int(kv["mask"]), "shfl": int(kv["shfl"]), "cs": int(kv["cs"]), "w": _nums(kv["w"]).reshape(256, 2)})
        elif head == "frame":
            plan["single_class"] = int(kv["single_class"])
            plan["f_init"] = _nums(kv["init"])
            plan["f_eweight"] = _nums(kv["eweight"])
        elif head == "f32":
            plan["f32"] = {"n": int(kv["n"]), "orig": _nums(kv["orig"], int), "init": _nums(kv["init"]), "eweight": _nums(kv["eweight"]), "segs": []}
        elif head == "f32seg":
            plan["f32"]["segs"].append({"partner": _nums(kv["partner"], int), "w": _nums(kv["w"]).reshape(32, 2)})
        elif head == "f16":
            plan["f16"] = {"n": int(kv["n"]), "scaled": int(kv["scaled"]), "const": int(kv["const"]), "lane": _nums(kv["lane"], int), "mask": _nums(kv["mask"], int),
                           "init": _nums(kv["init"]), "eweight": _nums(kv["eweight"]), "segs": [], "kappa": []}
            if "sweight" in kv:
                plan["f16"]["sweight"] = _nums(kv["sweight"])
            if "static" in kv:
                plan["f16"]["static"] = _nums(kv["static"], int)
                plan["f16"]["cstat"] = _nums(kv["cstat"])
                plan["f16"]["facrow"] = _nums(kv["facrow"], int)
        elif head == "f16seg":
            plan["f16"]["segs"].append(_nums(kv["tab"]).reshape(16, 4))
            if "kappa" in kv:
                plan["f16"]["kappa"].append(_nums(kv["kappa"]))
        elif head == "fclass":
            plan["fclasses"].append({"param": int(kv["param"]), "scale": float(kv["scale"]), "offset": float(kv["offset"])})
        elif head == "flabel":
            plan["flabels"][int(pos[0])] = _nums(kv["label"], int)
        elif head == "estage":
            plan["e_src"] = _nums(kv["src"], int)
            plan["e_w"] = _nums(kv["w"])
        elif head == "slot":
            plan["slots"][int(pos[0])] = {"nq": int(kv["nq"]), "seg": int(kv["seg"]), "tab": int(kv["tab"]),
                                          "diag_ok": _nums(kv["diag_ok"], int), "label": _nums(kv["label"], int)}
        elif head == "rot":
            plan["rots"].append({"param": int(kv["param"]), "scale": float(kv["scale"]), "offset": float(kv["offset"])})
    return plan


def init_state(n: int) -> np.ndarray:
    idx = np.arange(4 ** n)
    return np.where(idx & 0xAAAAAAAA, 0.0, 1.0)


def _apply_local(r, n, qubits, M):
    k = len(qubits)
    t = r.reshape((4,) * n)
    m = M.reshape((4,) * (2 * k))
    out = np.tensordot(m, t, axes=(list(range(k, 2 * k)), list(qubits)))
    return np.moveaxis(out, list(range(k)), list(qubits)).reshape(-1)


def run_blocks(plan, params=None, variants=None) -> np.ndarray:
    n = plan["n"]
    r = init_state(n)
    for b in plan["blocks"]:
        qs = [b["q0"]] if b["nq"] == 1 else [b["q0"], b["q1"]]
        dim = 4 ** b["nq"]
        if b["kind"] == 0:
            r = _apply_local(r, n, qs, b["M"].reshape(dim, dim))
        elif b["kind"] == 1:
            th = b["scale"] * params[b["param"]] + b["offset"]
            M = b["A"] + np.cos(th) * b["C"] + np.sin(th) * b["S"]
            r = _apply_local(r, n, qs, M.reshape(dim, dim))
        else:
            idx = 0 if variants is None else int(variants[b["slot"]])
            if idx == 0:
                continue
            tab = b["tab"].reshape(16, 4, 4)
            B = tab[idx] if b["nq"] == 1 else np.kron(tab[idx >> 4], tab[idx & 15])
            r = _apply_local(r, n, qs, b["cond"].reshape(dim, dim) @ B)
    return r


def energy(plan, r) -> float:
    return float(np.dot(plan["term_coef"], r[plan["term_idx"]]))


def run_devops(plan, params=None, variants=None) -> np.ndarray:
    """Interpret the device program (passes -> phases -> sweeps -> micro-ops) as dmx_tile_kernel executes it."""
    n = plan["n"]
    cf = plan["coefs"]
    pad = 2 if n == 1 else 0
    nbits = 2 * n + pad
    E = 1 << nbits
    idx = np.arange(E)
    r = np.zeros(E)
    r[np.arange(4 ** n) << pad] = init_state(n)

    def chain_matrix(ch):
        M = np.eye(4)
        for f in plan["factors"][ch["first"]:ch["first"] + ch["n"]]:
            if f["kind"] == 0:
                F = cf[f["coef"]:f["coef"] + 16].reshape(4, 4)
            else:
                rot = plan["rots"][f["rot"]]
                th = rot["scale"] * params[rot["param"]] + rot["offset"]
                acs = cf[f["coef"]:f["coef"] + 48].reshape(3, 4, 4)
                F = acs[0] + np.cos(th) * acs[1] + np.sin(th) * acs[2]
            M = F @ M
        return M

    for ps in plan["passes"]:
        tile = list(ps["tile"])
        k = len(tile)
        if n == 1:
            gpos = {2: 2, 0: 0}
        else:
            gpos = {2 * (k - 1 - i): 2 * (n - 1 - q) for i, q in enumerate(tile)}
        for ph in plan["phases"][ps["first_phase"]:ps["first_phase"] + ps["nphases"]]:
            mats = [chain_matrix(c) for c in plan["chains"][ph["first_chain"]:ph["first_chain"] + ph["nchains"]]]
            if plan["sweeps3"]:
                for sw in plan["sweeps3"][ph["first_sweep"]:ph["first_sweep"] + ph["nsweeps"]]:
                    r = _run_sweep3(plan, sw, mats, gpos, r, idx)
                continue
            for sw in plan["sweeps"][ph["first_sweep"]:ph["first_sweep"] + ph["nsweeps"]]:
                g0, g1 = gpos[sw["p0"]], gpos[sw["p1"]]
                c0, c1 = (idx >> g0) & 3, (idx >> g1) & 3
                loc = c0 * 4 + c1
                clear = idx & ~(3 << g0) & ~(3 << g1)

                def at(code):
                    return clear | ((code >> 2) << g0) | ((code & 3) << g1)

                # tile-local XOR masks -> global index masks (bit positions of the pair only)
                def gmask(m):
                    out = 0
                    for lp, gp in ((sw["p0"], g0), (sw["p1"], g1)):
                        out |= ((int(m) >> lp) & 3) << gp
                    return out

                def xor_index(masks):
                    x = np.zeros_like(idx)
                    for j in range(4):
                        x ^= np.where((loc >> j) & 1, gmask(masks[j]), 0)
                    return clear ^ x

                if sw["lperm"]:
                    v = cf[sw["lcoef"] + loc] * r[xor_index(sw["lmask"])]
                else:
                    assert np.array_equal(xor_index(sw["lmask"]), idx)
                    v = r.copy()

                def side(v, M, a_side, unital):
                    if unital:
                        assert M[0, 0] == 1 and not M[0, 1:].any() and not M[1:, 0].any()
                    new = np.zeros_like(v)
                    for b in range(4):
                        srcidx = (idx & ~(3 << (g0 if a_side else g1))) | (b << (g0 if a_side else g1))
                        new += M[(c0 if a_side else c1), b] * v[srcidx]
                    return new

                def dense(v, M):
                    new = np.zeros_like(v)
                    for b in range(16):
                        new += M[loc, b] * v[at(np.full_like(loc, b))]
                    return new

                for m in plan["mus"][sw["first_mu"]:sw["first_mu"] + sw["nmu"]]:
                    kind = m["kind"]
                    if kind in (0, 1):
                        v = side(v, mats[m["a0"]], kind == 0, m["a3"])
                    elif kind in (2, 3):
                        v = side(v, cf[m["a0"]:m["a0"] + 16].reshape(4, 4), kind == 2, m["a3"])
                    elif kind == 4:
                        v = v * cf[m["a0"] + c0]
                    elif kind == 5:
                        v = v * cf[m["a0"] + c1]
                    elif kind == 6:
                        v = v * cf[m["a0"] + loc]
                    elif kind == 7:
                        v = dense(v, cf[m["a0"]:m["a0"] + 256].reshape(16, 16))
                    else:
                        var = 0 if variants is None else int(variants[m["a0"]])
                        if var == 0:
                            continue
                        tab = cf[m["a1"]:m["a1"] + 256].reshape(16, 4, 4)
                        if kind == 8:
                            v = side(v, tab[var & 15], True, 0)
                        elif kind == 9:
                            v = side(v, tab[var & 15], False, 0)
                        else:
                            v = side(v, tab[var >> 4], True, 0)
                            v = side(v, tab[var & 15], False, 0)
                            if m["a3"] == 1:
                                v = v * cf[m["a2"] + loc]
                            else:
                                v = dense(v, cf[m["a2"]:m["a2"] + 256].reshape(16, 16))
                if sw["sperm"]:
                    out = np.zeros_like(v)
                    out[xor_index(sw["smask"])] = cf[sw["scoef"] + loc] * v
                    r = out
                else:
                    assert np.array_equal(xor_index(sw["smask"]), idx)
                    r = v
    return r[np.arange(4 ** n) << pad]


def _wide_side(sw, gpos, idx, loc, lp, masks, tmasks, vmasks, nvar, wide_side):
    """Global index and scale-table row of every element for one side of a (possibly wide) triple sweep: the element held by
    thread t in register e lives at tile index T(t) ^ X(e); T is the canonical placement of the thread's bits unless the side
    is wide (then XOR_j t_j * tmask[j]); the scale-table variant is sum_j parity(t & vmask[j]) << j."""
    tile_bits = sorted(gpos)                       # tile-local code positions (even), ascending
    k = len(tile_bits)
    # tile-local index of every global index
    a = np.zeros_like(idx)
    for l in tile_bits:
        a |= ((idx >> gpos[l]) & 3) << l
    outer = idx.copy()
    for l in tile_bits:
        outer &= ~(3 << gpos[l])
    # thread index: the non-triple tile bits, compacted in ascending order
    t = np.zeros_like(idx)
    j = 0
    for l in tile_bits:
        if l in lp:
            continue
        t |= ((a >> l) & 3) << j
        j += 2
    ntb = 2 * (k - 3)
    if wide_side:
        base = np.zeros_like(idx)
        for b in range(ntb):
            base ^= np.where((t >> b) & 1, int(tmasks[b]), 0)
    else:
        base = a.copy()
        for l in lp:
            base &= ~(3 << l)
    x = np.zeros_like(idx)
    for b in range(6):
        x ^= np.where((loc >> b) & 1, int(masks[b]), 0)
    tile_idx = base ^ x
    g = outer.copy()
    for l in tile_bits:
        g |= ((tile_idx >> l) & 3) << gpos[l]
    var = np.zeros_like(idx)
    if wide_side:
        for b in range(nvar):
            par = np.zeros_like(idx)
            m = int(vmasks[b])
            for bit in range(ntb):
                if (m >> bit) & 1:
                    par ^= (t >> bit) & 1
            var |= par << b
    return g, var


def _run_sweep3(plan, sw, mats, gpos, r, idx):
    """One triple sweep as dmx_stream3_kernel executes it: register e = 16*code(A) + 4*code(B) + code(C)."""
    cf = plan["coefs"]
    lp = (sw["p0"], sw["p1"], sw["p2"])
    if sw.get("wide"):
        return _run_sweep3_wide(plan, sw, mats, gpos, r, idx)
    g = [gpos[x] for x in lp]
    codes = [(idx >> gg) & 3 for gg in g]
    loc = codes[0] * 16 + codes[1] * 4 + codes[2]
    clear = idx & ~(3 << g[0]) & ~(3 << g[1]) & ~(3 << g[2])

    def gmask(m):
        out = 0
        for l, gg in zip(lp, g):
            out |= ((int(m) >> l) & 3) << gg
        return out

    def xor_index(masks):
        x = np.zeros_like(idx)
        for j in range(6):
            x ^= np.where((loc >> j) & 1, gmask(masks[j]), 0)
        return clear ^ x

    if sw["lperm"]:
        v = cf[sw["lcoef"] + loc] * r[xor_index(sw["lmask"])]
    else:
        assert np.array_equal(xor_index(sw["lmask"]), idx)
        v = r.copy()

    def one(v, M, slot, unital):
        if unital:
            assert M[0, 0] == 1 and not M[0, 1:].any() and not M[1:, 0].any()
        new = np.zeros_like(v)
        for b in range(4):
            new += M[codes[slot], b] * v[(idx & ~(3 << g[slot])) | (b << g[slot])]
        return new

    pairs = ((0, 1), (0, 2), (1, 2))
    for m in plan["mus"][sw["first_mu"]:sw["first_mu"] + sw["nmu"]]:
        kind = m["kind"]
        if kind == 16:
            M = mats[m["a0"]] if m["a2"] else cf[m["a0"]:m["a0"] + 16].reshape(4, 4)
            v = one(v, M, m["a1"], m["a3"])
        elif kind == 17:
            v = v * cf[m["a0"] + codes[m["a1"]]]
        elif kind == 18:
            s0, s1 = pairs[m["a1"]]
            v = v * cf[m["a0"] + codes[s0] * 4 + codes[s1]]
        elif kind == 19:
            s0, s1 = pairs[m["a1"]]
            M = cf[m["a0"]:m["a0"] + 256].reshape(16, 16)
            l2 = codes[s0] * 4 + codes[s1]
            new = np.zeros_like(v)
            for b in range(16):
                src = (idx & ~(3 << g[s0]) & ~(3 << g[s1])) | ((b >> 2) << g[s0]) | ((b & 3) << g[s1])
                new += M[l2, b] * v[src]
            v = new
        else:
            raise ValueError(f"unexpected micro-op kind {kind} in a triple sweep")
    if sw["sperm"]:
        out = np.zeros_like(v)
        out[xor_index(sw["smask"])] = cf[sw["scoef"] + loc] * v
        return out
    assert np.array_equal(xor_index(sw["smask"]), idx)
    return v


def run_segments(plan, params=None, variants=None):
    """Interpret the segment-kernel tables; returns the energy, or None when the variant row needs the
    general path (a non-diagonal correction op)."""
    cf = plan["coefs"]
    j = np.arange(256)
    r = np.where(j & 0xAA, 0.0, 1.0)
    act = []
    if variants is not None:
        for s, v in enumerate(variants):
            v = int(v)
            if v == 0 or s not in plan["slots"] or plan["slots"][s]["seg"] < 0:
                continue
            sl = plan["slots"][s]
            if not (int(sl["diag_ok"][v >> 5]) >> (v & 31)) & 1:
                return None
            act.append((s, v))
    nseg = plan["nseg"]
    for k in range(nseg + 1):
        for s, v in act:
            sl = plan["slots"][s]
            if sl["seg"] != k:
                continue
            lab = sl["label"]
            t = sl["tab"]
            if sl["nq"] == 1:
                f = cf[t + v * 4 + lab]
            else:
                f = cf[t + (v >> 4) * 4 + (lab >> 2)] * cf[t + (v & 15) * 4 + (lab & 3)] * cf[t + 64 + lab]
            r = r * f
        if k == nseg:
            break
        sg = plan["segs"][k]
        a = sg["w0"] * r[sg["src0"]]
        if sg["rot"] >= 0:
            rot = plan["rots"][sg["rot"]]
            th = rot["scale"] * params[rot["param"]] + rot["offset"]
            b = sg["w1"] * r[sg["src1"]]
            r = np.where(sg["active"] == 1, np.cos(th) * a + np.sin(th) * b, a)
        elif sg["rot"] == -2:
            r = np.where(sg["active"] == 1, a + sg["w1"] * r[sg["src1"]], a)
        else:
            r = a
    return float(np.dot(plan["e_w"], r[plan["e_src"]]))


def run_frame32(plan, params=None, variants=None):
    """Interpret the liveness-compacted frame tables as dmx_frame32_kernel does (array index = lane)."""
    f32 = plan.get("f32")
    if f32 is None:
        return None
    cf = plan["coefs"]
    orig = np.asarray(f32["orig"])
    r = f32["init"].copy()
    act = []
    if variants is not None:
        for s, v in enumerate(variants):
            v = int(v)
            if v == 0 or s not in plan["slots"] or plan["slots"][s]["seg"] < 0:
                continue
            sl = plan["slots"][s]
            if not (int(sl["diag_ok"][v >> 5]) >> (v & 31)) & 1:
                return None
            act.append((s, v))
    nseg = plan["nseg"]
    for k in range(nseg + 1):
        for s, v in act:
            sl = plan["slots"][s]
            if sl["seg"] != k:
                continue
            lab = np.where(orig >= 0, plan["flabels"][s][np.maximum(orig, 0)], 0)
            tb = sl["tab"]
            if sl["nq"] == 1:
                f = cf[tb + v * 4 + lab]
            else:
                f = cf[tb + (v >> 4) * 4 + (lab >> 2)] * cf[tb + (v & 15) * 4 + (lab & 3)] * cf[tb + 64 + lab]
            r = r * f
        if k == nseg:
            break
        fs, seg = plan["fsegs"][k], f32["segs"][k]
        other = r[np.asarray(seg["partner"]) & 31]
        if fs["cs"] >= 0:
            cl = plan["fclasses"][fs["cs"]]
            th = cl["scale"] * params[cl["param"]] + cl["offset"]
            cm, sm = np.cos(th), np.sin(th)
        else:
            cm = sm = 1.0
        w0, w1 = seg["w"][:, 0], seg["w"][:, 1]
        cme = np.where((np.asarray(seg["partner"]) & 0x80) != 0, cm, 1.0)     # bit 7: the rotation moves this coefficient
        r = cme * (w0 * r) + sm * (w1 * other)
    return float(np.dot(f32["eweight"], r))


def run_frame16t(plan, params):
    """Interpret the thread-per-circuit tables as dmx_frame16t_kernel does: 16 slots, slot i pairs with i ^ mask[k]."""
    f16 = plan.get("f16")
    if f16 is None:
        return None
    cl = plan["fclasses"][0]
    th = cl["scale"] * params[cl["param"]] + cl["offset"]
    c, s = np.cos(th), np.sin(th)
    r = f16["init"].copy()
    idx = np.arange(16)
    for k, tab in enumerate(f16["segs"]):
        m = int(f16["mask"][k])
        own = tab[:, 0] * c + tab[:, 1]
        r = own * r + (tab[:, 2] * s) * r[idx ^ m]
    return float(np.dot(f16["eweight"][:16], r) + f16["eweight"][16])


def run_frame16t_scaled(plan, params):
    """The scaled form of dmx_frame16t_kernel: rho_i' = cos rho_i + kappa_i sin rho_j, kappa stored per pair in loop order."""
    f16 = plan.get("f16")
    if f16 is None or not f16["scaled"]:
        return None
    cl = plan["fclasses"][0]
    th = cl["scale"] * params[cl["param"]] + cl["offset"]
    c, s = np.cos(th), np.sin(th)
    r = f16["init"].copy()
    for k, kap in enumerate(f16["kappa"]):
        m = int(f16["mask"][k])
        new = r.copy()
        pair = 0
        for i in range(16):
            j = i ^ m
            if j > i:
                new[i] = kap[2 * pair] * (s * r[j]) + c * r[i]
                new[j] = kap[2 * pair + 1] * (s * r[i]) + c * r[j]
                pair += 1
        r = new
    return float(np.dot(f16["sweight"][:16], r) + f16["sweight"][16])


def run_frame16v(plan, variants):
    """The constant-coefficient thread-per-row form (dmx_frame16v_kernel): rho_i' = rho_i + kappa_i rho_j per segment, the row's
    diagonal corrections applied by stage in (stage, slot) order to the 16 slots and to the untouched coefficients."""
    f16 = plan.get("f16")
    if f16 is None or not f16["const"]:
        return None
    orig = np.asarray(plan["f32"]["orig"])
    cf = plan["coefs"]
    lanes, stat = [int(x) for x in f16["lane"]], [int(x) for x in f16["static"]]
    ents = []
    for s, v in enumerate(variants if variants is not None else []):
        v = int(v)
        if v == 0 or s not in plan["slots"] or plan["slots"][s]["seg"] < 0:
            continue
        sl = plan["slots"][s]
        if not (int(sl["diag_ok"][v >> 5]) >> (v & 31)) & 1:
            return None
        assert f16["facrow"][s] >= 0
        ents.append((int(sl["seg"]), s, v))
    ents.sort()

    def fac(s, v, lane):
        sl = plan["slots"][s]
        lab = int(plan["flabels"][s][orig[lane]])
        tb = sl["tab"]
        if sl["nq"] == 1:
            return cf[tb + v * 4 + lab]
        return cf[tb + (v >> 4) * 4 + (lab >> 2)] * cf[tb + (v & 15) * 4 + (lab & 3)] * cf[tb + 64 + lab]

    r = f16["init"].copy()
    F = np.ones(len(stat))
    nseg = plan["nseg"]
    for k in range(nseg + 1):
        for seg, s, v in ents:
            if seg != k:
                continue
            for i, lane in enumerate(lanes):
                if lane >= 0:
                    r[i] *= fac(s, v, lane)
            for j, lane in enumerate(stat):
                F[j] *= fac(s, v, lane)
        if k == nseg:
            break
        m, kap = int(f16["mask"][k]), f16["kappa"][k]
        new, pair = r.copy(), 0
        for i in range(16):
            j = i ^ m
            if j > i:
                new[i] = kap[2 * pair] * r[j] + r[i]
                new[j] = kap[2 * pair + 1] * r[i] + r[j]
                pair += 1
        r = new
    return float(np.dot(f16["sweight"][:16], r) + np.dot(f16["cstat"][:len(stat)], F))


def run_frame(plan, params=None, variants=None):
    """Interpret the fixed-frame tables exactly as dmx_frame_kernel does (thread t = array index t)."""
    cf = plan["coefs"]
    t = np.arange(256)
    r = plan["f_init"].copy()
    act = []
    if variants is not None:
        for s, v in enumerate(variants):
            v = int(v)
            if v == 0 or s not in plan["slots"] or plan["slots"][s]["seg"] < 0:
                continue
            sl = plan["slots"][s]
            if not (int(sl["diag_ok"][v >> 5]) >> (v & 31)) & 1:
                return None
            act.append((s, v))
    nseg = plan["nseg"]
    for k in range(nseg + 1):
        for s, v in act:
            sl = plan["slots"][s]
            if sl["seg"] != k:
                continue
            lab = plan["flabels"][s]
            tb = sl["tab"]
            if sl["nq"] == 1:
                f = cf[tb + v * 4 + lab]
            else:
                f = cf[tb + (v >> 4) * 4 + (lab >> 2)] * cf[tb + (v & 15) * 4 + (lab & 3)] * cf[tb + 64 + lab]
            r = r * f
        if k == nseg:
            break
        fs = plan["fsegs"][k]
        if fs["shfl"]:
            assert fs["mask"] < 32
        other = r[t ^ fs["mask"]]
        if fs["cs"] >= 0:
            cl = plan["fclasses"][fs["cs"]]
            th = cl["scale"] * params[cl["param"]] + cl["offset"]
            cm, sm = np.cos(th), np.sin(th)
        else:
            cm = sm = 1.0
        w0, w1 = fs["w"][:, 0], fs["w"][:, 1]
        cme = np.where(w1 != 0.0, cm, 1.0)
        r = cme * (w0 * r) + sm * (w1 * other)
    return float(np.dot(plan["f_eweight"], r))


def _run_sweep3_wide(plan, sw, mats, gpos, r, idx):
    """A wide sweep = wide load (gather + variant scale table), the local body on canonical coordinates, wide store."""
    cf = plan["coefs"]
    lp = (sw["p0"], sw["p1"], sw["p2"])
    g = [gpos[x] for x in lp]
    codes = [(idx >> gg) & 3 for gg in g]
    loc = codes[0] * 16 + codes[1] * 4 + codes[2]
    if sw["lperm"]:
        src, var = _wide_side(sw, gpos, idx, loc, lp, sw["lmask"], sw["ltmask"], sw["lvmask"], sw["lnvar"], sw["wide"] & 1)
        assert len(np.unique(src)) == len(src)
        v = cf[sw["lcoef"] + var * 64 + loc] * r[src]
    else:
        v = r.copy()
    body = dict(sw)
    body.update(lperm=0, sperm=0, wide=0, lmask=[1 << (lp[2] + 0), 1 << (lp[2] + 1), 1 << (lp[1] + 0), 1 << (lp[1] + 1), 1 << (lp[0] + 0), 1 << (lp[0] + 1)])
    body["smask"] = body["lmask"]
    v = _run_sweep3(plan, body, mats, gpos, v, idx)
    if sw["sperm"]:
        dst, var = _wide_side(sw, gpos, idx, loc, lp, sw["smask"], sw["stmask"], sw["svmask"], sw["snvar"], sw["wide"] & 2)
        assert len(np.unique(dst)) == len(dst)
        out = np.zeros_like(v)
        out[dst] = cf[sw["scoef"] + var * 64 + loc] * v
        return out
    return v
