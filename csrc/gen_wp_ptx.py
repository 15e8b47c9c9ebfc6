#!/usr/bin/env python
"""Generates csrc/agg_wp_ptx.inc: the PTX statement of the two hot blocks of k_agg_group_wp (agg_wp.cuh) for E entry slots per
lane.  nvcc turns every `bool` that guards an inline-asm access into SEL + ISETP, which more than doubled the instruction count
of the C++ statement of these steps; writing the blocks as one asm each keeps the guards in predicate registers.
    python csrc/gen_wp_ptx.py > csrc/agg_wp_ptx.inc
"""
E = 8


def q(s):
    return '        "' + s + '\\n"'


def probe():
    out = []
    out.append("// first probe of the CTA's key table for %d entries: one LDS.128 of the key's bucket {key0,id0,key1,id1} per active entry." % E)
    out.append("// id[j] = the dense id when the bucket holds the key with a published id; returns the mask of active entries that still need")
    out.append("// the slow path (displaced key, first sighting, id in flight).")
    out.append("__device__ __forceinline__ uint32_t wp_probe%d(uint32_t act, const uint32_t (&key)[%d], uint32_t kt_addr, uint32_t hash_shift, uint32_t (&id)[%d]) {" % (E, E, E))
    out.append("    uint32_t need;")
    out.append('    asm volatile("{\\n"')
    preds = ", ".join("p%d" % j for j in range(E))
    out.append(q(" .reg .pred %s, h, g;" % preds))
    regs = ", ".join("a%d, ka%d, ia%d, kb%d, ib%d" % (j, j, j, j, j) for j in range(E))
    out.append(q(" .reg .b32 t, %s;" % regs))
    # operands: %0 need, %1..%E id, %(E+1) act, %(E+2)..%(2E+1) key, %(2E+2) kt_addr, %(2E+3) shift
    o_act, o_key, o_kt, o_sh = E + 1, E + 2, 2 * E + 2, 2 * E + 3
    for j in range(E):
        out.append(q(" and.b32 t, %%%d, %d; setp.ne.u32 p%d, t, 0;" % (o_act, 1 << j, j)))
    for j in range(E):
        out.append(q(" mul.lo.u32 a%d, %%%d, 0x9E3779B1; shr.u32 a%d, a%d, %%%d; shl.b32 a%d, a%d, 4; add.u32 a%d, a%d, %%%d;"
                     % (j, o_key + j, j, j, o_sh, j, j, j, j, o_kt)))
    for j in range(E):
        out.append(q(" @p%d ld.volatile.shared.v4.u32 {ka%d, ia%d, kb%d, ib%d}, [a%d];" % (j, j, j, j, j, j)))
    out.append(q(" mov.u32 %0, 0;"))
    for j in range(E):
        k = "%%%d" % (o_key + j)
        out.append(q(" setp.eq.u32 h, ka%d, %s; setp.lt.and.u32 h, ia%d, 0xFFFFFFF0, h; setp.eq.u32 g, kb%d, %s; setp.lt.and.u32 g, ib%d, 0xFFFFFFF0, g;"
                     % (j, k, j, j, k, j)))
        out.append(q(" selp.u32 %%%d, ia%d, ib%d, h; or.pred h, h, g; and.pred h, p%d, !h; @h or.b32 %%0, %%0, %d;" % (1 + j, j, j, j, 1 << j)))
    out.append('        "}\\n"')
    out.append("        : \"=r\"(need), " + ", ".join('"=r"(id[%d])' % j for j in range(E)))
    out.append("        : \"r\"(act), " + ", ".join('"r"(key[%d])' % j for j in range(E)) + ', "r"(kt_addr), "r"(hash_shift)')
    out.append('        : "memory");')
    out.append("    return need;")
    out.append("}")
    return out


def rounds(nv):
    """nv = 1 or 2 double sums per group.  Operands: %0 pm (in/out), %1..%E addr64 (cnt address | sums address << 32),
    %(E+1) tag word of entry 0 ((1 << 8) + lane), then E doubles of column a (and E of column b)."""
    out = []
    name = "wp_round%d_f64x%d" % (E, nv)
    out.append("// one arbitration round for %d entries (%d double sum%s per group): pending entries read their count word, write" % (E, nv, "s" if nv > 1 else ""))
    out.append("// (count + 1 | tag), read it back; the entries whose word survived add their values.  Returns `pm` without the winners.")
    sig = "__device__ __forceinline__ uint32_t %s(uint32_t pm, const uint64_t (&ad)[%d], uint32_t tag0, const uint64_t (&va)[%d]" % (name, E, E)
    if nv == 2:
        sig += ", const uint64_t (&vb)[%d]" % E
    out.append(sig + ") {")
    out.append('    asm volatile("{\\n"')
    out.append(q(" .reg .pred %s, %s;" % (", ".join("p%d" % j for j in range(E)), ", ".join("w%d" % j for j in range(E)))))
    out.append(q(" .reg .b32 t, %s;" % ", ".join("ca%d, sa%d, c%d, b%d, g%d" % (j, j, j, j, j) for j in range(E))))
    fr = ", ".join("x%d, s%d" % (j, j) for j in range(E))
    if nv == 2:
        fr += ", " + ", ".join("y%d, u%d" % (j, j) for j in range(E))
    out.append(q(" .reg .f64 %s;" % fr))
    o_ad, o_tag, o_va, o_vb = 1, E + 1, E + 2, 2 * E + 2
    for j in range(E):
        out.append(q(" and.b32 t, %%0, %d; setp.ne.u32 p%d, t, 0; mov.b64 {ca%d, sa%d}, %%%d; add.u32 g%d, %%%d, %d;" % (1 << j, j, j, j, o_ad + j, j, o_tag, 32 * j)))
    for j in range(E):
        out.append(q(" @p%d ld.volatile.shared.u32 c%d, [ca%d];" % (j, j, j)))
    out.append(q(" bar.warp.sync 0xffffffff;"))
    for j in range(E):
        out.append(q(" and.b32 c%d, c%d, 0xFFFFFF00; add.u32 c%d, c%d, g%d;" % (j, j, j, j, j)))
    for j in range(E):
        out.append(q(" @p%d st.volatile.shared.u32 [ca%d], c%d;" % (j, j, j)))
    out.append(q(" bar.warp.sync 0xffffffff;"))
    for j in range(E):
        out.append(q(" @p%d ld.volatile.shared.u32 b%d, [ca%d];" % (j, j, j)))
    for j in range(E):
        out.append(q(" setp.eq.and.u32 w%d, b%d, c%d, p%d;" % (j, j, j, j)))
    for j in range(E):
        if nv == 2:
            out.append(q(" @w%d ld.volatile.shared.v2.f64 {x%d, y%d}, [sa%d];" % (j, j, j, j)))
        else:
            out.append(q(" @w%d ld.volatile.shared.f64 x%d, [sa%d];" % (j, j, j)))
    for j in range(E):
        if nv == 2:
            out.append(q(" add.f64 s%d, x%d, %%%d; add.f64 u%d, y%d, %%%d;" % (j, j, o_va + j, j, j, o_vb + j)))
        else:
            out.append(q(" add.f64 s%d, x%d, %%%d;" % (j, j, o_va + j)))
    for j in range(E):
        if nv == 2:
            out.append(q(" @w%d st.volatile.shared.v2.f64 [sa%d], {s%d, u%d};" % (j, j, j, j)))
        else:
            out.append(q(" @w%d st.volatile.shared.f64 [sa%d], s%d;" % (j, j, j)))
    for j in range(E):
        out.append(q(" @w%d and.b32 %%0, %%0, %d;" % (j, (~(1 << j)) & ((1 << E) - 1))))
    out.append(q(" bar.warp.sync 0xffffffff;"))
    out.append('        "}\\n"')
    out.append('        : "+r"(pm)')
    ins = ", ".join('"l"(ad[%d])' % j for j in range(E)) + ', "r"(tag0), ' + ", ".join('"d"(bits_f64(va[%d]))' % j for j in range(E))
    if nv == 2:
        ins += ", " + ", ".join('"d"(bits_f64(vb[%d]))' % j for j in range(E))
    out.append("        : " + ins)
    out.append('        : "memory");')
    out.append("    return pm;")
    out.append("}")
    return out


print("// GENERATED by csrc/gen_wp_ptx.py — do not edit by hand.")
print("constexpr int WP_E = %d;   // entry slots per lane and iteration (two row quads)" % E)
for block in (probe(), rounds(2), rounds(1)):
    print("\n".join(block))
    print()
