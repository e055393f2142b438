"""A small AArch64 (A64) user-mode interpreter: the integer, load/store, branch, scalar floating-point and the few
Advanced-SIMD forms that clang emits for Primer3's thal.c / thal_main.c.  Every instruction word is decoded once into a Python
closure (cached by address).  An encoding that is not handled raises `Unknown` with the word and its address - nothing is
guessed.  Written from the Arm ARM encodings; validated end to end by reproducing, from the reference's own executable,
the five `ntthal` output blocks that delta_g.rs:197-230 holds (tests/golden/ntthal_delta_g_rs.json).

Test infrastructure only (see emu.py)."""
import math
import struct
from fractions import Fraction

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1
M128 = (1 << 128) - 1


class Unknown(Exception):
    pass


class Halt(Exception):
    pass


def sx(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def d2b(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def b2d(b):
    return struct.unpack("<d", struct.pack("<Q", b & M64))[0]


def s2b(x):
    try:
        return struct.unpack("<I", struct.pack("<f", x))[0]
    except OverflowError:
        return 0x7F800000 if x > 0 else 0xFF800000


def b2s(b):
    return struct.unpack("<f", struct.pack("<I", b & M32))[0]


def fdiv(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


def fma(a, b, c):
    """Correctly rounded a*b+c (Python 3.12 has no math.fma)."""
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return a * b + c
    r = Fraction(a) * Fraction(b) + Fraction(c)
    if r == 0:
        p = a * b
        return p + c  # signed zero by the ordinary rules
    try:
        return float(r)
    except OverflowError:
        return math.inf if r > 0 else -math.inf


def decode_bitmasks(n, imms, immr, datasize):
    v = (n << 6) | ((~imms) & 0x3F)
    ln = v.bit_length() - 1
    if ln < 1:
        raise Unknown("reserved bitmask immediate")
    esize = 1 << ln
    levels = esize - 1
    s, r = imms & levels, immr & levels
    d = (s - r) & levels
    emask = (1 << esize) - 1
    welem = (1 << (s + 1)) - 1
    telem = (1 << (d + 1)) - 1
    welem = ((welem >> r) | (welem << (esize - r))) & emask if r else welem
    w = t = 0
    for i in range(datasize // esize):
        w |= welem << (i * esize)
        t |= telem << (i * esize)
    return w, t


class CPU:
    def __init__(self, mem, base):
        self.mem = mem
        self.base = base
        self.size = len(mem)
        self.x = [0] * 32        # x[31] = SP
        self.v = [0] * 32        # 128-bit
        self.n = self.z = self.c = self.vf = 0
        self.cache = {}
        self.hooks = {}          # address -> python callable(cpu) (imports)
        self.icount = 0

    # ---- memory
    def rd(self, addr, n):
        o = addr - self.base
        if o < 0 or o + n > self.size:
            raise MemoryError("read %#x" % addr)
        return int.from_bytes(self.mem[o: o + n], "little")

    def wr(self, addr, n, val):
        o = addr - self.base
        if o < 0 or o + n > self.size:
            raise MemoryError("write %#x" % addr)
        self.mem[o: o + n] = (val & ((1 << (8 * n)) - 1)).to_bytes(n, "little")

    def cstr(self, addr):
        o = addr - self.base
        e = self.mem.index(b"\0", o)
        return bytes(self.mem[o:e])

    # ---- flags
    def cond(self, c):
        b = c >> 1
        if b == 0:
            r = self.z
        elif b == 1:
            r = self.c
        elif b == 2:
            r = self.n
        elif b == 3:
            r = self.vf
        elif b == 4:
            r = self.c and not self.z
        elif b == 5:
            r = self.n == self.vf
        elif b == 6:
            r = (self.n == self.vf) and not self.z
        else:
            return True
        return (not r) if (c & 1) else bool(r)

    def addc(self, a, b, cin, bits, setflags):
        mask = (1 << bits) - 1
        u = a + b + cin
        r = u & mask
        if setflags:
            self.n = r >> (bits - 1)
            self.z = 1 if r == 0 else 0
            self.c = 1 if u > mask else 0
            sa, sb = a >> (bits - 1), b >> (bits - 1)
            self.vf = 1 if (sa == sb and self.n != sa) else 0
        return r

    def fcmp_flags(self, a, b):
        if a != a or b != b:
            self.n, self.z, self.c, self.vf = 0, 0, 1, 1
        elif a == b:
            self.n, self.z, self.c, self.vf = 0, 1, 1, 0
        elif a < b:
            self.n, self.z, self.c, self.vf = 1, 0, 0, 0
        else:
            self.n, self.z, self.c, self.vf = 0, 0, 1, 0

    # ---- run
    def run(self, pc, stop, limit=None):
        cache, hooks, mem, base = self.cache, self.hooks, self.mem, self.base
        x = self.x
        n = 0
        try:
          while pc != stop:
            f = cache.get(pc)
            if f is None:
                h = hooks.get(pc)
                if h is not None:
                    h(self)
                    pc = x[30]
                    continue
                w = int.from_bytes(mem[pc - base: pc - base + 4], "little")
                try:
                    f = decode(self, w, pc)
                except Unknown as e:
                    raise Unknown("%s: word %08x at %#x" % (e, w, pc)) from None
                cache[pc] = f
            r = f()
            pc = pc + 4 if r is None else r
            n += 1
            if limit is not None and n > limit:
                raise Halt("instruction limit")
        finally:
            self.icount += n
        return n


def decode(cpu, w, pc):
    x, v = cpu.x, cpu.v
    rd_, wr_ = cpu.rd, cpu.wr
    op0 = (w >> 25) & 0xF
    Rd = w & 31
    Rn = (w >> 5) & 31
    Rm = (w >> 16) & 31
    sf = w >> 31
    bits = 64 if sf else 32
    mask = M64 if sf else M32

    def setr(r, val):          # write Xd/Wd, register 31 = ZR
        if r != 31:
            x[r] = val
    # ------------------------------------------------------------------ data processing, immediate
    if op0 in (8, 9):
        grp = (w >> 23) & 7
        if (w >> 24) & 0x1F == 0x10:   # ADR / ADRP
            imm = sx(((w >> 5) & 0x7FFFF) << 2 | ((w >> 29) & 3), 21)
            val = ((pc & ~0xFFF) + (imm << 12)) & M64 if sf else (pc + imm) & M64

            def f():
                if Rd != 31:
                    x[Rd] = val
            return f
        if grp == 2:                 # ADD/SUB immediate
            imm = (w >> 10) & 0xFFF
            if (w >> 22) & 1:
                imm <<= 12
            sub, S = (w >> 30) & 1, (w >> 29) & 1
            b = ((~imm) & mask) if sub else imm
            cin = 1 if sub else 0
            if S:
                def f():
                    r = cpu.addc(x[Rn] & mask, b, cin, bits, True)   # Rn=31 is SP
                    if Rd != 31:
                        x[Rd] = r
            else:
                def f():
                    x[Rd] = (x[Rn] + b + cin) & mask                  # both SP-capable
            return f
        if grp == 4:                 # logical immediate
            opc = (w >> 29) & 3
            N = (w >> 22) & 1
            if N and not sf:
                raise Unknown("logical imm N=1 sf=0")
            imm, _ = decode_bitmasks(N, (w >> 10) & 0x3F, (w >> 16) & 0x3F, bits)

            def f():
                a = (x[Rn] if Rn != 31 else 0) & mask
                if opc == 0:
                    x[Rd] = a & imm              # Rd=31 is SP
                elif opc == 1:
                    x[Rd] = a | imm
                elif opc == 2:
                    x[Rd] = a ^ imm
                else:
                    r = a & imm
                    cpu.n, cpu.z, cpu.c, cpu.vf = r >> (bits - 1), int(r == 0), 0, 0
                    if Rd != 31:
                        x[Rd] = r
            return f
        if grp == 5:                 # MOVN/MOVZ/MOVK
            opc = (w >> 29) & 3
            sh = ((w >> 21) & 3) * 16
            imm = (w >> 5) & 0xFFFF
            if opc == 0:
                val = (~(imm << sh)) & mask
            elif opc == 2:
                val = (imm << sh) & mask
            elif opc == 3:
                keep = (~(0xFFFF << sh)) & mask
                ins = imm << sh

                def f():
                    if Rd != 31:
                        x[Rd] = (x[Rd] & keep) | ins
                return f
            else:
                raise Unknown("move wide opc=1")

            def f():
                if Rd != 31:
                    x[Rd] = val
            return f
        if grp == 6:                 # bitfield
            opc = (w >> 29) & 3
            N = (w >> 22) & 1
            immr, imms = (w >> 16) & 0x3F, (w >> 10) & 0x3F
            wm, tm = decode_bitmasks(N, imms, immr, bits)
            R, S = immr, imms

            def ror(a):
                return ((a >> R) | (a << (bits - R))) & mask if R else a
            if opc == 0:             # SBFM
                def f():
                    src = (x[Rn] if Rn != 31 else 0) & mask
                    bot = ror(src) & wm
                    top = mask if (src >> S) & 1 else 0
                    setr(Rd, (top & ~tm & mask) | (bot & tm))
            elif opc == 1:           # BFM
                def f():
                    src = (x[Rn] if Rn != 31 else 0) & mask
                    dst = (x[Rd] if Rd != 31 else 0) & mask
                    bot = (dst & ~wm & mask) | (ror(src) & wm)
                    setr(Rd, (dst & ~tm & mask) | (bot & tm))
            elif opc == 2:           # UBFM
                def f():
                    src = (x[Rn] if Rn != 31 else 0) & mask
                    setr(Rd, ror(src) & wm & tm)
            else:
                raise Unknown("bitfield opc=3")
            return f
        if grp == 7:                 # EXTR
            lsb = (w >> 10) & 0x3F

            def f():
                hi = (x[Rn] if Rn != 31 else 0) & mask
                lo = (x[Rm] if Rm != 31 else 0) & mask
                setr(Rd, (((hi << bits) | lo) >> lsb) & mask)
            return f
        raise Unknown("dp-imm")
    # ------------------------------------------------------------------ branches, system
    if op0 in (10, 11):
        top6 = w >> 26
        if top6 in (0x05, 0x25):     # B / BL
            tgt = (pc + (sx(w & 0x3FFFFFF, 26) << 2)) & M64
            if top6 == 0x25:
                ret = pc + 4

                def f():
                    x[30] = ret
                    return tgt
            else:
                def f():
                    return tgt
            return f
        if (w >> 24) == 0x54:        # B.cond
            tgt = (pc + (sx((w >> 5) & 0x7FFFF, 19) << 2)) & M64
            c = w & 15
            cond = cpu.cond

            def f():
                return tgt if cond(c) else None
            return f
        if (w >> 25) & 0x3F == 0x1A:  # CBZ / CBNZ
            tgt = (pc + (sx((w >> 5) & 0x7FFFF, 19) << 2)) & M64
            nz = (w >> 24) & 1
            Rt = Rd

            def f():
                val = (x[Rt] if Rt != 31 else 0) & mask
                return tgt if (val != 0) == bool(nz) else None
            return f
        if (w >> 25) & 0x3F == 0x1B:  # TBZ / TBNZ
            tgt = (pc + (sx((w >> 5) & 0x3FFF, 14) << 2)) & M64
            nz = (w >> 24) & 1
            bit = ((w >> 31) << 5) | ((w >> 19) & 31)
            Rt = Rd

            def f():
                val = ((x[Rt] if Rt != 31 else 0) >> bit) & 1
                return tgt if val == nz else None
            return f
        if (w & 0xFFFFFC1F) == 0xD61F0000:   # BR
            def f():
                return x[Rn]
            return f
        if (w & 0xFFFFFC1F) == 0xD63F0000:   # BLR
            ret = pc + 4

            def f():
                t = x[Rn]
                x[30] = ret
                return t
            return f
        if (w & 0xFFFFFC1F) == 0xD65F0000:   # RET
            def f():
                return x[Rn]
            return f
        if (w & 0xFFFFF01F) == 0xD503201F:   # hints (NOP, BTI, PAC hints)
            return lambda: None
        if (w & 0xFFE0001F) == 0xD4200000:   # BRK
            def f():
                raise Halt("brk #%d at %#x" % ((w >> 5) & 0xFFFF, pc))
            return f
        raise Unknown("branch/system")
    # ------------------------------------------------------------------ loads and stores
    if op0 & 5 == 4:
        V = (w >> 26) & 1
        Rt = Rd
        if (w >> 27) & 7 == 3 and (w >> 24) & 3 == 0:   # LDR literal
            opc = w >> 30
            addr = (pc + (sx((w >> 5) & 0x7FFFF, 19) << 2)) & M64
            if V:
                nb = 4 << opc

                def f():
                    v[Rt] = rd_(addr, nb)
            elif opc == 0:
                def f():
                    setr(Rt, rd_(addr, 4))
            elif opc == 1:
                def f():
                    setr(Rt, rd_(addr, 8))
            elif opc == 2:
                def f():
                    setr(Rt, sx(rd_(addr, 4), 32) & M64)
            else:
                return lambda: None   # PRFM
            return f
        if (w >> 27) & 7 == 5:       # LDP / STP
            opc = w >> 30
            mode = (w >> 23) & 3
            L = (w >> 22) & 1
            Rt2 = (w >> 10) & 31
            if V:
                nb = 4 << opc
            else:
                if opc == 0:
                    nb = 4
                elif opc == 2:
                    nb = 8
                elif opc == 1 and L:
                    nb = 4     # LDPSW
                else:
                    raise Unknown("ldp/stp opc")
            off = sx((w >> 15) & 0x7F, 7) * nb
            signed = (not V) and opc == 1

            def f():
                b = x[Rn]
                a = b if mode == 1 else (b + off) & M64
                if L:
                    a0, a1 = rd_(a, nb), rd_(a + nb, nb)
                    if V:
                        v[Rt], v[Rt2] = a0, a1
                    else:
                        if signed:
                            a0, a1 = sx(a0, 32) & M64, sx(a1, 32) & M64
                        setr(Rt, a0)
                        setr(Rt2, a1)
                else:
                    if V:
                        wr_(a, nb, v[Rt])
                        wr_(a + nb, nb, v[Rt2])
                    else:
                        wr_(a, nb, x[Rt] if Rt != 31 else 0)
                        wr_(a + nb, nb, x[Rt2] if Rt2 != 31 else 0)
                if mode in (1, 3):
                    x[Rn] = (b + off) & M64
            if mode == 0:
                raise Unknown("ldnp/stnp")
            return f
        if (w >> 27) & 7 == 7:       # LDR/STR single register
            size = w >> 30
            opc = (w >> 22) & 3
            if V:
                scale = size | ((opc >> 1) << 2)
                if scale > 4:
                    raise Unknown("simd ldr size")
                nb = 1 << scale
                load = opc & 1
                kind = 0
            else:
                scale = size
                nb = 1 << size
                if opc == 0:
                    load, kind = 0, 0
                elif opc == 1:
                    load, kind = 1, 0
                elif size == 3 and opc == 2:
                    load, kind = 2, 0        # PRFM
                elif opc == 2:
                    load, kind = 1, 64
                else:
                    if size >= 2:
                        raise Unknown("ldrs opc=3 size>=2")
                    load, kind = 1, 32

            def access(a):
                if load == 2:
                    return
                if load:
                    val = rd_(a, nb)
                    if V:
                        v[Rt] = val
                    else:
                        if kind == 64:
                            val = sx(val, 8 * nb) & M64
                        elif kind == 32:
                            val = sx(val, 8 * nb) & M32
                        if Rt != 31:
                            x[Rt] = val
                else:
                    wr_(a, nb, v[Rt] if V else (x[Rt] if Rt != 31 else 0))
            if (w >> 24) & 1:        # unsigned immediate
                off = ((w >> 10) & 0xFFF) << scale

                def f():
                    access((x[Rn] + off) & M64)
                return f
            if (w >> 21) & 1:        # register offset
                if (w >> 10) & 3 != 2:
                    raise Unknown("atomic / pac load")
                option = (w >> 13) & 7
                S = (w >> 12) & 1
                sh = scale if S else 0
                if option == 3:       # LSL / UXTX

                    def f():
                        access((x[Rn] + ((x[Rm] if Rm != 31 else 0) << sh)) & M64)
                elif option == 2:     # UXTW
                    def f():
                        access((x[Rn] + (((x[Rm] if Rm != 31 else 0) & M32) << sh)) & M64)
                elif option == 6:     # SXTW
                    def f():
                        access((x[Rn] + (sx((x[Rm] if Rm != 31 else 0), 32) << sh)) & M64)
                elif option == 7:
                    def f():
                        access((x[Rn] + (sx((x[Rm] if Rm != 31 else 0), 64) << sh)) & M64)
                else:
                    raise Unknown("ldr reg option")
                return f
            mode = (w >> 10) & 3
            off = sx((w >> 12) & 0x1FF, 9)
            if mode == 0:            # unscaled
                def f():
                    access((x[Rn] + off) & M64)
            elif mode == 1:          # post-index
                def f():
                    b = x[Rn]
                    access(b)
                    x[Rn] = (b + off) & M64
            elif mode == 3:          # pre-index
                def f():
                    a = (x[Rn] + off) & M64
                    access(a)
                    x[Rn] = a
            else:
                raise Unknown("ldtr/sttr")
            return f
        if (w & 0xBFFF0000) in (0x0C400000, 0x0C000000) or (w & 0xBFE00000) in (0x0CC00000, 0x0C800000):
            # LD1/ST1 multiple structures (1..4 registers), optional post-index
            opcode = (w >> 12) & 0xF
            nreg = {7: 1, 10: 2, 6: 3, 2: 4}.get(opcode)
            if nreg is None:
                raise Unknown("ldN/stN structure form")
            Q = (w >> 30) & 1
            L = (w >> 22) & 1
            post = (w >> 23) & 1
            nb = 16 if Q else 8

            def f():
                a = x[Rn]
                for i in range(nreg):
                    r = (Rt + i) & 31
                    if L:
                        v[r] = rd_(a + i * nb, nb)
                    else:
                        wr_(a + i * nb, nb, v[r])
                if post:
                    inc = nreg * nb if Rm == 31 else x[Rm]
                    x[Rn] = (a + inc) & M64
            return f
        raise Unknown("load/store")
    # ------------------------------------------------------------------ data processing, register
    if op0 & 7 == 5:
        op1 = (w >> 28) & 1
        op2 = (w >> 21) & 0xF
        if not op1 and not (op2 & 8):                 # logical shifted register
            opc = (w >> 29) & 3
            N = (w >> 21) & 1
            st = (w >> 22) & 3
            amt = (w >> 10) & 0x3F

            def f():
                a = (x[Rn] if Rn != 31 else 0) & mask
                b = (x[Rm] if Rm != 31 else 0) & mask
                if amt:
                    if st == 0:
                        b = (b << amt) & mask
                    elif st == 1:
                        b >>= amt
                    elif st == 2:
                        b = (sx(b, bits) >> amt) & mask
                    else:
                        b = ((b >> amt) | (b << (bits - amt))) & mask
                if N:
                    b = (~b) & mask
                if opc == 0:
                    r = a & b
                elif opc == 1:
                    r = a | b
                elif opc == 2:
                    r = a ^ b
                else:
                    r = a & b
                    cpu.n, cpu.z, cpu.c, cpu.vf = r >> (bits - 1), int(r == 0), 0, 0
                if Rd != 31:
                    x[Rd] = r
            return f
        if not op1 and (op2 & 9) == 8:                # add/sub shifted register
            sub, S = (w >> 30) & 1, (w >> 29) & 1
            st = (w >> 22) & 3
            amt = (w >> 10) & 0x3F

            def f():
                a = (x[Rn] if Rn != 31 else 0) & mask
                b = (x[Rm] if Rm != 31 else 0) & mask
                if amt:
                    if st == 0:
                        b = (b << amt) & mask
                    elif st == 1:
                        b >>= amt
                    else:
                        b = (sx(b, bits) >> amt) & mask
                if sub:
                    r = cpu.addc(a, (~b) & mask, 1, bits, S)
                else:
                    r = cpu.addc(a, b, 0, bits, S)
                if Rd != 31:
                    x[Rd] = r
            return f
        if not op1 and (op2 & 9) == 9:                # add/sub extended register
            sub, S = (w >> 30) & 1, (w >> 29) & 1
            option = (w >> 13) & 7
            sh = (w >> 10) & 7

            def f():
                a = x[Rn] & mask                      # SP
                b = x[Rm] if Rm != 31 else 0
                ln = 8 << (option & 3)
                b = (sx(b, ln) if option & 4 else b & ((1 << ln) - 1))
                b = (b << sh) & mask
                if sub:
                    r = cpu.addc(a, (~b) & mask, 1, bits, S)
                else:
                    r = cpu.addc(a, b, 0, bits, S)
                if S:
                    if Rd != 31:
                        x[Rd] = r
                else:
                    x[Rd] = r                         # SP
            return f
        if op1:
            if op2 == 0 and (w >> 10) & 0x3F == 0:    # ADC / SBC
                sub, S = (w >> 30) & 1, (w >> 29) & 1

                def f():
                    a = (x[Rn] if Rn != 31 else 0) & mask
                    b = (x[Rm] if Rm != 31 else 0) & mask
                    if sub:
                        b = (~b) & mask
                    setr(Rd, cpu.addc(a, b, cpu.c, bits, S))
                return f
            if op2 == 2:                              # CCMP / CCMN
                sub = (w >> 30) & 1
                isimm = (w >> 11) & 1
                c = (w >> 12) & 15
                nzcv = w & 15

                def f():
                    if cpu.cond(c):
                        a = (x[Rn] if Rn != 31 else 0) & mask
                        b = Rm if isimm else ((x[Rm] if Rm != 31 else 0) & mask)
                        if sub:
                            cpu.addc(a, (~b) & mask, 1, bits, True)
                        else:
                            cpu.addc(a, b, 0, bits, True)
                    else:
                        cpu.n, cpu.z, cpu.c, cpu.vf = (nzcv >> 3) & 1, (nzcv >> 2) & 1, (nzcv >> 1) & 1, nzcv & 1
                return f
            if op2 == 4:                              # CSEL / CSINC / CSINV / CSNEG
                c = (w >> 12) & 15
                op = (w >> 30) & 1
                o2 = (w >> 10) & 1

                def f():
                    if cpu.cond(c):
                        r = (x[Rn] if Rn != 31 else 0) & mask
                    else:
                        r = (x[Rm] if Rm != 31 else 0) & mask
                        if op:
                            r = (~r) & mask
                        if o2:
                            r = (r + 1) & mask
                    if Rd != 31:
                        x[Rd] = r
                return f
            if op2 & 8:                               # 3-source
                op31 = (w >> 21) & 7
                o0 = (w >> 15) & 1
                Ra = (w >> 10) & 31
                if op31 == 0:
                    def f():
                        a = (x[Rn] if Rn != 31 else 0) & mask
                        b = (x[Rm] if Rm != 31 else 0) & mask
                        c_ = (x[Ra] if Ra != 31 else 0) & mask
                        setr(Rd, (c_ - a * b if o0 else c_ + a * b) & mask)
                elif op31 in (1, 5) and sf:           # SMADDL/SMSUBL, UMADDL/UMSUBL
                    sg = op31 == 1

                    def f():
                        a = (x[Rn] if Rn != 31 else 0) & M32
                        b = (x[Rm] if Rm != 31 else 0) & M32
                        if sg:
                            a, b = sx(a, 32), sx(b, 32)
                        c_ = (x[Ra] if Ra != 31 else 0)
                        setr(Rd, (c_ - a * b if o0 else c_ + a * b) & M64)
                elif op31 in (2, 6) and sf:           # SMULH / UMULH
                    sg = op31 == 2

                    def f():
                        a = (x[Rn] if Rn != 31 else 0)
                        b = (x[Rm] if Rm != 31 else 0)
                        if sg:
                            a, b = sx(a, 64), sx(b, 64)
                        setr(Rd, ((a * b) >> 64) & M64)
                else:
                    raise Unknown("dp3")
                return f
            if op2 == 6:
                opcode = (w >> 10) & 0x3F
                if (w >> 30) & 1:                     # 1-source
                    def f():
                        a = (x[Rn] if Rn != 31 else 0) & mask
                        if opcode == 0:
                            r = int(format(a, "0%db" % bits)[::-1], 2)
                        elif opcode == 4:
                            r = bits - a.bit_length()
                        elif opcode == 5:
                            s = a >> (bits - 1)
                            t = (a ^ (mask if s else 0)) & mask
                            r = bits - 1 - t.bit_length()
                        elif opcode == 1:
                            r = int.from_bytes(b"".join(a.to_bytes(bits // 8, "little")[i:i + 2][::-1] for i in range(0, bits // 8, 2)), "little")
                        elif (opcode == 2 and not sf) or (opcode == 3 and sf):
                            r = int.from_bytes(a.to_bytes(bits // 8, "little"), "big")
                        elif opcode == 2 and sf:
                            bb = a.to_bytes(8, "little")
                            r = int.from_bytes(bb[0:4][::-1] + bb[4:8][::-1], "little")
                        else:
                            raise Unknown("dp1 opcode")
                        setr(Rd, r)
                    if opcode > 5:
                        raise Unknown("dp1 opcode")
                    return f
                if opcode in (2, 3):                  # UDIV / SDIV
                    sg = opcode == 3

                    def f():
                        a = (x[Rn] if Rn != 31 else 0) & mask
                        b = (x[Rm] if Rm != 31 else 0) & mask
                        if b == 0:
                            r = 0
                        elif sg:
                            a, b = sx(a, bits), sx(b, bits)
                            q = abs(a) // abs(b)
                            r = (-q if (a < 0) != (b < 0) else q) & mask
                        else:
                            r = a // b
                        setr(Rd, r)
                    return f
                if opcode in (8, 9, 10, 11):          # LSLV / LSRV / ASRV / RORV
                    def f():
                        a = (x[Rn] if Rn != 31 else 0) & mask
                        s = (x[Rm] if Rm != 31 else 0) % bits
                        if opcode == 8:
                            r = (a << s) & mask
                        elif opcode == 9:
                            r = a >> s
                        elif opcode == 10:
                            r = (sx(a, bits) >> s) & mask
                        else:
                            r = ((a >> s) | (a << (bits - s))) & mask if s else a
                        setr(Rd, r)
                    return f
                raise Unknown("dp2 opcode")
        raise Unknown("dp-reg")
    # ------------------------------------------------------------------ scalar FP and SIMD
    if op0 & 7 == 7:
        return decode_fp(cpu, w, pc)
    raise Unknown("top-level group")


def decode_fp(cpu, w, pc):
    x, v = cpu.x, cpu.v
    Rd = w & 31
    Rn = (w >> 5) & 31
    Rm = (w >> 16) & 31
    if (w & 0x5F200000) == 0x1E200000:               # scalar FP (M=0,S=0)
        ftype = (w >> 22) & 3
        if ftype not in (0, 1):
            raise Unknown("fp16")
        dbl = ftype == 1
        fmask = M64 if dbl else M32
        get = (lambda r: b2d(v[r])) if dbl else (lambda r: b2s(v[r]))
        put = d2b if dbl else s2b
        sf = w >> 31
        if (w >> 10) & 0x3F == 0:                    # FP <-> integer conversions
            rmode, opcode = (w >> 19) & 3, (w >> 16) & 7
            ibits = 64 if sf else 32
            imask = M64 if sf else M32
            if opcode in (2, 3) and rmode == 0:      # SCVTF / UCVTF
                sg = opcode == 2

                def f():
                    a = (x[Rn] if Rn != 31 else 0) & imask
                    if sg:
                        a = sx(a, ibits)
                    v[Rd] = put(float(a))
                return f
            if opcode in (0, 1):                     # FCVT{N,P,M,Z}{S,U}
                sg = opcode == 0

                def f():
                    a = get(Rn)
                    if a != a:
                        r = 0
                    elif math.isinf(a):
                        r = (1 << 70) if a > 0 else -(1 << 70)
                    elif rmode == 3:
                        r = math.trunc(a)
                    elif rmode == 2:
                        r = math.floor(a)
                    elif rmode == 1:
                        r = math.ceil(a)
                    else:
                        r = round(a)
                    if sg:
                        lo, hi = -(1 << (ibits - 1)), (1 << (ibits - 1)) - 1
                    else:
                        lo, hi = 0, imask
                    r = min(max(r, lo), hi)
                    if Rd != 31:
                        x[Rd] = r & imask
                return f
            if opcode == 6 and rmode == 0:           # FMOV Xd/Wd <- Dn/Sn
                def f():
                    if Rd != 31:
                        x[Rd] = v[Rn] & fmask
                return f
            if opcode == 7 and rmode == 0:           # FMOV Dd/Sd <- Xn/Wn
                def f():
                    v[Rd] = (x[Rn] if Rn != 31 else 0) & fmask
                return f
            raise Unknown("fp<->int conversion")
        if sf:
            raise Unknown("fp sf=1")
        low = (w >> 10) & 3
        if low == 2:                                 # 2-source
            opcode = (w >> 12) & 15
            if opcode == 0:
                op = lambda a, b: a * b
            elif opcode == 1:
                op = fdiv
            elif opcode == 2:
                op = lambda a, b: a + b
            elif opcode == 3:
                op = lambda a, b: a - b
            elif opcode in (4, 6):
                op = lambda a, b: (a if a != a else b if b != b else max(a, b)) if opcode == 4 else (b if a != a else a if b != b else max(a, b))
            elif opcode in (5, 7):
                op = lambda a, b: (a if a != a else b if b != b else min(a, b)) if opcode == 5 else (b if a != a else a if b != b else min(a, b))
            elif opcode == 8:
                op = lambda a, b: -(a * b)
            else:
                raise Unknown("fp 2-source opcode")

            def f():
                try:
                    r = op(get(Rn), get(Rm))
                except OverflowError:
                    r = math.inf
                v[Rd] = put(r)
            return f
        if low == 1:                                 # FCCMP
            c = (w >> 12) & 15
            nzcv = w & 15

            def f():
                if cpu.cond(c):
                    cpu.fcmp_flags(get(Rn), get(Rm))
                else:
                    cpu.n, cpu.z, cpu.c, cpu.vf = (nzcv >> 3) & 1, (nzcv >> 2) & 1, (nzcv >> 1) & 1, nzcv & 1
            return f
        if low == 3:                                 # FCSEL
            c = (w >> 12) & 15

            def f():
                v[Rd] = (v[Rn] if cpu.cond(c) else v[Rm]) & fmask
            return f
        # low == 0
        if (w >> 10) & 0xF == 8:                     # FCMP / FCMPE
            zero = (w >> 3) & 1

            def f():
                cpu.fcmp_flags(get(Rn), 0.0 if zero else get(Rm))
            return f
        if (w >> 10) & 7 == 4:                       # FMOV immediate
            imm8 = (w >> 13) & 0xFF
            sign = imm8 >> 7
            # VFPExpandImm: exponent = NOT(b6) : b6 replicated : b5 b4  ->  2^(1..4) for b6 = 0, 2^(-3..0) for b6 = 1
            e = (((imm8 >> 4) & 3) + 1) if not (imm8 >> 6) & 1 else (((imm8 >> 4) & 3) - 3)
            val = (16 + (imm8 & 15)) / 16.0 * (2.0 ** e)
            if sign:
                val = -val
            bitsv = put(val)

            def f():
                v[Rd] = bitsv
            return f
        if (w >> 10) & 0x1F == 0x10:                 # 1-source
            opcode = (w >> 15) & 0x3F
            if opcode == 0:
                def f():
                    v[Rd] = v[Rn] & fmask
            elif opcode == 1:
                def f():
                    v[Rd] = (v[Rn] & fmask) & (fmask >> 1)
            elif opcode == 2:
                def f():
                    v[Rd] = (v[Rn] & fmask) ^ ((fmask >> 1) + 1)
            elif opcode == 3:
                def f():
                    a = get(Rn)
                    v[Rd] = put(math.sqrt(a) if a >= 0 else math.nan)
            elif opcode == 5 and not dbl:            # FCVT S -> D
                def f():
                    v[Rd] = d2b(b2s(v[Rn]))
            elif opcode == 4 and dbl:                # FCVT D -> S
                def f():
                    v[Rd] = s2b(b2d(v[Rn]))
            elif opcode in (8, 9, 10, 11, 12, 14, 15):   # FRINT N,P,M,Z,A,X,I
                def f():
                    a = get(Rn)
                    if a != a or math.isinf(a) or a == 0.0:
                        r = a
                    elif opcode == 9:
                        r = float(math.ceil(a))
                    elif opcode == 10:
                        r = float(math.floor(a))
                    elif opcode == 11:
                        r = float(math.trunc(a))
                    elif opcode == 12:
                        r = float(math.floor(abs(a) + 0.5)) * (1 if a > 0 else -1)
                    else:
                        r = float(round(a))
                    if r == 0.0:
                        r = math.copysign(0.0, a)
                    v[Rd] = put(r)
            else:
                raise Unknown("fp 1-source opcode")
            return f
        raise Unknown("scalar fp")
    if (w & 0xFF000000) == 0x1F000000:               # FMADD / FMSUB / FNMADD / FNMSUB
        ftype = (w >> 22) & 3
        if ftype != 1:
            raise Unknown("fp 3-source single")
        o1, o0 = (w >> 21) & 1, (w >> 15) & 1
        Ra = (w >> 10) & 31

        def f():
            a, b, c = b2d(v[Rn]), b2d(v[Rm]), b2d(v[Ra])
            if o1:
                c = -c
            if o0 != o1:
                a = -a
            v[Rd] = d2b(fma(a, b, c))
        return f
    # ---- Advanced SIMD, only the forms met in the binary
    Q = (w >> 30) & 1
    regmask = M128 if Q else M64
    if (w & 0x9FF80C00) == 0x0F000400:               # modified immediate (MOVI / MVNI / FMOV vector)
        op = (w >> 29) & 1
        cmode = (w >> 12) & 15
        imm8 = ((w >> 16) & 7) << 5 | ((w >> 5) & 31)
        if cmode == 14 and op == 1:                  # MOVI 64-bit byte mask
            imm64 = 0
            for i in range(8):
                if (imm8 >> i) & 1:
                    imm64 |= 0xFF << (8 * i)
        elif cmode == 14 and op == 0:                # MOVI 8-bit
            imm64 = int.from_bytes(bytes([imm8]) * 8, "little")
        elif cmode & 9 == 0:                         # 32-bit shifted (MOVI / MVNI)
            e = imm8 << (8 * (cmode >> 1))
            if op:
                e = (~e) & M32
            imm64 = e | (e << 32)
        elif cmode & 13 == 8:                        # 16-bit shifted
            e = imm8 << (8 * ((cmode >> 1) & 1))
            if op:
                e = (~e) & 0xFFFF
            imm64 = e * 0x0001000100010001
        elif cmode == 15 and op == 1 and Q:        # FMOV Vd.2D, #imm
            e = (((imm8 >> 4) & 3) + 1) if not (imm8 >> 6) & 1 else (((imm8 >> 4) & 3) - 3)
            val_ = (16 + (imm8 & 15)) / 16.0 * (2.0 ** e)
            imm64 = d2b(-val_ if imm8 >> 7 else val_)
        else:
            raise Unknown("simd modified immediate cmode=%d op=%d" % (cmode, op))
        val = (imm64 | (imm64 << 64)) & regmask

        def f():
            v[Rd] = val
        return f
    if (w & 0xBFE0FC00) == 0x0E000C00:               # DUP (general)
        imm5 = (w >> 16) & 31
        sz = (imm5 & -imm5).bit_length() - 1
        if sz > 3:
            raise Unknown("dup imm5")
        eb = 8 << sz
        cnt = (128 if Q else 64) // eb

        def f():
            e = (x[Rn] if Rn != 31 else 0) & ((1 << eb) - 1)
            r = 0
            for i in range(cnt):
                r |= e << (i * eb)
            v[Rd] = r
        return f
    if (w & 0xBFE0FC00) == 0x0E000400:               # DUP (element)
        imm5 = (w >> 16) & 31
        sz = (imm5 & -imm5).bit_length() - 1
        eb = 8 << sz
        idx = imm5 >> (sz + 1)
        cnt = (128 if Q else 64) // eb

        def f():
            e = (v[Rn] >> (idx * eb)) & ((1 << eb) - 1)
            r = 0
            for i in range(cnt):
                r |= e << (i * eb)
            v[Rd] = r
        return f
    if (w & 0xFFE0FC00) == 0x5E000400:               # DUP / MOV scalar element
        imm5 = (w >> 16) & 31
        sz = (imm5 & -imm5).bit_length() - 1
        eb = 8 << sz
        idx = imm5 >> (sz + 1)

        def f():
            v[Rd] = (v[Rn] >> (idx * eb)) & ((1 << eb) - 1)
        return f
    if (w & 0xFFE0FC00) == 0x4E001C00:               # INS (general)
        imm5 = (w >> 16) & 31
        sz = (imm5 & -imm5).bit_length() - 1
        eb = 8 << sz
        idx = imm5 >> (sz + 1)
        em = (1 << eb) - 1

        def f():
            e = (x[Rn] if Rn != 31 else 0) & em
            v[Rd] = (v[Rd] & ~(em << (idx * eb)) & M128) | (e << (idx * eb))
        return f
    if (w & 0xFFE08400) == 0x6E000400:               # INS (element)
        imm5 = (w >> 16) & 31
        imm4 = (w >> 11) & 15
        sz = (imm5 & -imm5).bit_length() - 1
        eb = 8 << sz
        di = imm5 >> (sz + 1)
        si = imm4 >> sz
        em = (1 << eb) - 1

        def f():
            e = (v[Rn] >> (si * eb)) & em
            v[Rd] = (v[Rd] & ~(em << (di * eb)) & M128) | (e << (di * eb))
        return f
    if (w & 0xBFE0FC00) == 0x0E003C00:               # UMOV / MOV to general
        imm5 = (w >> 16) & 31
        sz = (imm5 & -imm5).bit_length() - 1
        eb = 8 << sz
        idx = imm5 >> (sz + 1)

        def f():
            if Rd != 31:
                x[Rd] = (v[Rn] >> (idx * eb)) & ((1 << eb) - 1)
        return f
    if (w & 0xFFE0FC00) == 0x7EE0D400:               # FABD Dd, Dn, Dm
        def f():
            v[Rd] = d2b(abs(b2d(v[Rn]) - b2d(v[Rm])))
        return f
    if (w & 0xFFFFFC00) == 0x7E70D800:               # FADDP Dd, Vn.2D
        def f():
            v[Rd] = d2b(b2d(v[Rn]) + b2d(v[Rn] >> 64))
        return f
    if (w & 0x9F3E0C00) == 0x0E200800:               # two-register miscellaneous
        U = (w >> 29) & 1
        size = (w >> 22) & 3
        opcode = (w >> 12) & 31
        eb = 8 << size
        em = (1 << eb) - 1
        cnt = (128 if Q else 64) // eb
        if opcode not in (8, 9, 10, 11) or (U and opcode == 10):
            raise Unknown("simd two-reg misc opcode %d U=%d" % (opcode, U))

        def f():
            a = v[Rn]
            r = 0
            for i in range(cnt):
                e = sx((a >> (i * eb)) & em, eb)
                if opcode == 8:
                    t = (e >= 0) if U else (e > 0)
                elif opcode == 9:
                    t = (e <= 0) if U else (e == 0)
                elif opcode == 10:
                    t = e < 0
                else:
                    r |= ((-e if U else abs(e)) & em) << (i * eb)
                    continue
                if t:
                    r |= em << (i * eb)
            v[Rd] = r
        return f
    if (w & 0xBF3FFC00) == 0x0E31B800:               # ADDV
        size = (w >> 22) & 3
        eb = 8 << size
        em = (1 << eb) - 1
        cnt = (128 if Q else 64) // eb

        def f():
            a = v[Rn]
            v[Rd] = sum((a >> (i * eb)) & em for i in range(cnt)) & em
        return f
    if (w & 0x9F200400) == 0x0E200400:               # three same
        U = (w >> 29) & 1
        size = (w >> 22) & 3
        opcode = (w >> 11) & 31
        if opcode == 3:                              # AND/BIC/ORR/ORN/EOR/BSL/BIT/BIF
            def f():
                a, b, d = v[Rn] & regmask, v[Rm] & regmask, v[Rd] & regmask
                if not U:
                    r = (a & b, a & ~b, a | b, a | ~b)[size]
                else:
                    r = (a ^ b, b ^ ((b ^ a) & d), d ^ ((d ^ a) & b), d ^ ((d ^ a) & ~b))[size]
                v[Rd] = r & regmask
            return f
        if opcode in (16, 6, 12, 13, 17, 7):         # ADD/SUB, CMGT/CMHI, MAX/MIN, CMEQ/CMTST, CMGE/CMHS
            eb = 8 << size
            em = (1 << eb) - 1
            cnt = (128 if Q else 64) // eb

            def f():
                a, b = v[Rn], v[Rm]
                r = 0
                for i in range(cnt):
                    ea, ebv = (a >> (i * eb)) & em, (b >> (i * eb)) & em
                    if opcode == 16:
                        e = (ea - ebv if U else ea + ebv) & em
                    else:
                        if not U:
                            sa, sb = sx(ea, eb), sx(ebv, eb)
                        else:
                            sa, sb = ea, ebv
                        if opcode == 6:
                            e = em if sa > sb else 0
                        elif opcode == 7:
                            e = em if sa >= sb else 0
                        elif opcode == 12:
                            e = max(sa, sb) & em
                        elif opcode == 13:
                            e = min(sa, sb) & em
                        else:
                            e = (em if ea == ebv else 0) if U else (em if ea & ebv else 0)
                    r |= e << (i * eb)
                v[Rd] = r
            return f
        raise Unknown("simd three-same opcode %d" % opcode)
    raise Unknown("fp/simd")
