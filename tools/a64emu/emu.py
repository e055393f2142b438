"""Run the reference's own Primer3 2.6.1 `ntthal` executable (od-msspe/bin/ntthal, Mach-O arm64 - the program
delta_g.rs:90-108 spawns) inside this x86 image: the image is mapped by macho.py, its instructions are interpreted by
cpu.py, and the ~40 libSystem imports it binds are served by the Python functions below (stdio on in-memory buffers,
a bump allocator, libm's log, the ctype tables).  What comes out is the executable's own stdout for a given command
line, i.e. REFERENCE OUTPUT, which no restatement in this repository had a hand in.

Test infrastructure: used offline by tools/gen_ntthal_emulated_golden.py to write tests/golden/ntthal_emulated.json.
Nothing under open-msspe-design_b200/ imports it; the GPU box never runs it (there is no /root/reference there).

    python tools/a64emu/emu.py -a HAIRPIN -s1 CCGCAGTAAGCTGCGG
"""
import math
import os
import re
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cpu import CPU, Halt, M64, b2d, d2b, sx   # noqa: E402
from macho import Image                         # noqa: E402

NTTHAL = "/root/reference/od-msspe/bin/ntthal"
MEM_SIZE = 48 << 20
HOOK_BASE = 0xF00000000000
STOP = 0xF0000FFF0000
_FMT = re.compile(rb"%([-+ #0]*)(\*|\d+)?(?:\.(\*|\d+))?(hh|h|ll|l|L|z|j|t)?([diouxXcsfFeEgGp%])")


class Exit(Exception):
    def __init__(self, code):
        self.code = code


class Process:
    def __init__(self, path=NTTHAL, file_root=None, mem_size=MEM_SIZE):
        self.im = im = Image(path)
        self.base = im.base
        self.mem_size = mem_size
        self.mem = bytearray(mem_size)
        binds = im.map_into(self.mem, self.base)
        self.cpu = cpu = CPU(self.mem, self.base)
        image_end = max(v + sz for _n, v, sz, _fo, _fs in im.segments if _n != "__PAGEZERO") - self.base
        self.heap = self.base + ((image_end + 0xFFFFF) & ~0xFFFFF)
        self.heap_end = self.base + mem_size - (4 << 20)
        self.sizes = {}
        self.out = {1: bytearray(), 2: bytearray()}
        self.stdin = b""
        self.stdin_pos = 0
        self.files = {}
        self.file_root = file_root
        self.calls = {}
        data = {"___stack_chk_guard": self._cell(0x5AFE5AFE5AFE5AFE), "___stdinp": self._cell(0x10),
                "___stdoutp": self._cell(0x11), "___stderrp": self._cell(0x12), "__DefaultRuneLocale": self._rune_locale(),
                "_optarg": self._cell(0), "_opterr": self._cell(0), "_optind": self._cell(1)}
        for i, (addr, name, addend) in enumerate(binds):
            if name in data:
                cpu.wr(addr, 8, data[name] + addend)
                continue
            fn = getattr(self, "imp" + name, None)
            if fn is None:
                raise NotImplementedError("import %s" % name)
            h = HOOK_BASE + 16 * i
            cpu.hooks[h] = self._wrap(name, fn)
            cpu.wr(addr, 8, h)

    # ---- plumbing
    def _wrap(self, name, fn):
        def h(cpu):
            self.calls[name] = self.calls.get(name, 0) + 1
            r = fn()
            if r is not None:
                cpu.x[0] = r & M64
        return h

    def _cell(self, val):
        a = self.alloc(8)
        self.cpu.wr(a, 8, val)
        return a

    def alloc(self, n):
        a = (self.heap + 15) & ~15
        self.heap = a + max(n, 1)
        if self.heap > self.heap_end:
            raise MemoryError("emulated heap exhausted")
        self.sizes[a] = n
        return a

    def _rune_locale(self):
        a = self._rl = self.alloc(60 + 3 * 1024)
        for c in range(256):
            ch = chr(c)
            t = 0
            if c < 128:
                if ch.isalpha():
                    t |= 0x100
                if c < 32 or c == 127:
                    t |= 0x200
                if ch.isdigit():
                    t |= 0x400 | (c - 48)
                if 32 < c < 127:
                    t |= 0x800
                if ch.islower():
                    t |= 0x1000
                if 32 < c < 127 and not ch.isalnum():
                    t |= 0x2000
                if ch in " \t\n\v\f\r":
                    t |= 0x4000
                if ch.isupper():
                    t |= 0x8000
                if ch in "0123456789abcdefABCDEF":
                    t |= 0x10000
                if ch in " \t":
                    t |= 0x20000
                if 32 <= c < 127:
                    t |= 0x40000
            self.cpu.wr(a + 60 + 4 * c, 4, t)
            self.cpu.wr(a + 60 + 1024 + 4 * c, 4, ord(ch.lower()) if c < 128 else c)
            self.cpu.wr(a + 60 + 2048 + 4 * c, 4, ord(ch.upper()) if c < 128 else c)
        return a

    def _bytes(self, addr, n):
        o = addr - self.base
        return bytes(self.mem[o: o + n])

    def _put(self, addr, data):
        o = addr - self.base
        if o < 0 or o + len(data) > self.mem_size:
            raise MemoryError("write %#x" % addr)
        self.mem[o: o + len(data)] = data

    def _format(self, fmt, ap):
        """C printf formatting; variadic arguments are 8-byte stack slots starting at `ap` (Apple arm64 ABI)."""
        cpu = self.cpu
        state = [ap]

        def nxt():
            val = cpu.rd(state[0], 8)
            state[0] += 8
            return val

        def sub(m):
            flags, width, prec, _ln, conv = m.groups()
            if conv == b"%":
                return b"%"
            flags = flags.decode()
            if width == b"*":
                wv = sx(nxt(), 32)
                if wv < 0:
                    flags += "-"
                    wv = -wv
                width = str(wv)
            else:
                width = width.decode() if width else ""
            if prec == b"*":
                prec = str(max(sx(nxt(), 32), 0))
            else:
                prec = prec.decode() if prec is not None else None
            spec = "%" + flags + width + ("." + prec if prec is not None else "")
            c = conv.decode()
            if c in "di":
                raw = nxt()
                val = sx(raw, 64) if _ln in (b"l", b"ll", b"z", b"j", b"t") else sx(raw, 32)
                return (spec + "d").encode() % val
            if c in "ouxX":
                raw = nxt()
                val = raw if _ln in (b"l", b"ll", b"z", b"j", b"t") else raw & 0xFFFFFFFF
                return (spec + c).encode() % val
            if c == "c":
                return (spec + "c").encode() % bytes([nxt() & 0xFF])
            if c == "s":
                s = cpu.cstr(nxt())
                return (spec + "s").encode() % s
            if c == "p":
                return b"0x%x" % nxt()
            val = b2d(nxt())
            return (spec + c).encode() % val
        return _FMT.sub(sub, fmt)

    def _write(self, stream, data):
        fd = {0x11: 1, 0x12: 2}.get(stream)
        if fd is None:
            raise IOError("write to FILE* %#x" % stream)
        self.out[fd] += data

    # ---- imports (arguments in x0.., d0..; result returned -> x0)
    def imp___error(self):
        return self._cell(0)

    def imp___maskrune(self):
        x = self.cpu.x
        c = x[0] & M64
        return self.cpu.rd(self._rl + 60 + 4 * c, 4) & x[1] if c < 256 else 0

    def imp___toupper(self):
        c = self.cpu.x[0] & 0xFFFFFFFF
        return ord(chr(c).upper()) if c < 128 else c

    def imp___stack_chk_fail(self):
        raise Halt("stack check failed")

    def imp_bsearch(self):
        """bsearch(key, base, nel, width, compar): compar is emulated code."""
        x = self.cpu.x
        key, base, nel, width, compar = x[0], x[1], x[2], x[3], x[4]
        saved = (x[30], x[19:30], x[31])
        lo, hi, found = 0, nel, 0
        while lo < hi:
            mid = (lo + hi) // 2
            el = base + mid * width
            x[0], x[1], x[30] = key, el, STOP
            self.cpu.run(compar, STOP)
            r = sx(x[0], 32)
            if r == 0:
                found = el
                break
            if r < 0:
                hi = mid
            else:
                lo = mid + 1
        x[30], x[31] = saved[0], saved[2]
        return found

    def imp_bzero(self):
        x = self.cpu.x
        self._put(x[0], bytes(x[1]))
        return None

    def imp_calloc(self):
        x = self.cpu.x
        return self.alloc(x[0] * x[1])          # the arena is zero and never reused

    def imp_malloc(self):
        return self.alloc(self.cpu.x[0])

    def imp_realloc(self):
        x = self.cpu.x
        old, n = x[0], x[1]
        a = self.alloc(n)
        if old:
            self._put(a, self._bytes(old, min(self.sizes.get(old, 0), n)))
        return a

    def imp_free(self):
        return None

    def imp_exit(self):
        raise Exit(sx(self.cpu.x[0], 32))

    def imp_fopen(self):
        x = self.cpu.x
        path = self.cpu.cstr(x[0]).decode()
        if self.file_root is None:
            return 0
        real = os.path.join(self.file_root, os.path.basename(path))
        if not os.path.isfile(real):
            return 0
        h = 0x100 + len(self.files)
        self.files[h] = [open(real, "rb").read(), 0, False]
        return h

    def imp_fclose(self):
        return 0

    def _stream(self, h):
        if h == 0x10:
            return None
        return self.files[h]

    def imp_feof(self):
        h = self.cpu.x[0]
        if h == 0x10:
            return int(self.stdin_eof)
        return int(self.files[h][2])

    stdin_eof = False

    def imp_fflush(self):
        return 0

    def imp_fgetc(self):
        h = self.cpu.x[0]
        if h == 0x10:
            if self.stdin_pos >= len(self.stdin):
                self.stdin_eof = True
                return 0xFFFFFFFF
            c = self.stdin[self.stdin_pos]
            self.stdin_pos += 1
            return c
        f = self.files[h]
        if f[1] >= len(f[0]):
            f[2] = True
            return 0xFFFFFFFF
        c = f[0][f[1]]
        f[1] += 1
        return c

    def imp_fgets(self):
        x = self.cpu.x
        buf, n, h = x[0], sx(x[1], 32), x[2]
        if h == 0x10:
            data, pos = self.stdin, self.stdin_pos
        else:
            data, pos = self.files[h][0], self.files[h][1]
        if pos >= len(data):
            if h == 0x10:
                self.stdin_eof = True
            else:
                self.files[h][2] = True
            return 0
        e = data.find(b"\n", pos)
        e = len(data) if e < 0 else e + 1
        e = min(e, pos + n - 1)
        self._put(buf, data[pos:e] + b"\0")
        if h == 0x10:
            self.stdin_pos = e
        else:
            self.files[h][1] = e
        return buf

    def imp_fprintf(self):
        x = self.cpu.x
        self._write(x[0], self._format(self.cpu.cstr(x[1]), x[31]))
        return 0

    def imp_printf(self):
        x = self.cpu.x
        self._write(0x11, self._format(self.cpu.cstr(x[0]), x[31]))
        return 0

    def imp_snprintf(self):
        x = self.cpu.x
        s = self._format(self.cpu.cstr(x[2]), x[31])
        n = x[1]
        if n:
            self._put(x[0], s[: n - 1] + b"\0")
        return len(s)

    def imp_fputs(self):
        x = self.cpu.x
        self._write(x[1], self.cpu.cstr(x[0]))
        return 0

    def imp_puts(self):
        self._write(0x11, self.cpu.cstr(self.cpu.x[0]) + b"\n")
        return 0

    def imp_putchar(self):
        self._write(0x11, bytes([self.cpu.x[0] & 0xFF]))
        return self.cpu.x[0] & 0xFF

    def imp_log(self):
        a = b2d(self.cpu.v[0])
        if a > 0:
            r = math.log(a)
        elif a == 0:
            r = -math.inf
        else:
            r = math.nan
        self.cpu.v[0] = d2b(r)
        return None

    def imp_setjmp(self):
        """The callee-saved state, keyed by the jmp_buf address (the buffer itself is not written)."""
        c = self.cpu
        self.jmp = getattr(self, "jmp", {})
        self.jmp[c.x[0]] = (list(c.x[19:32]), list(c.v[8:16]))
        return 0

    def imp_longjmp(self):
        c = self.cpu
        st = getattr(self, "jmp", {}).get(c.x[0])
        if st is None:
            raise Halt("longjmp without setjmp: " + self.out[2].decode(errors="replace"))
        val = c.x[1] & 0xFFFFFFFF
        c.x[19:32] = st[0]
        c.v[8:16] = st[1]
        return val or 1          # resumes at the saved x30: the hook dispatcher continues at x[30]

    def imp_memset(self):
        x = self.cpu.x
        self._put(x[0], bytes([x[1] & 0xFF]) * x[2])
        return x[0]

    def imp_memset_pattern16(self):
        x = self.cpu.x
        pat = self._bytes(x[1], 16)
        n = x[2]
        self._put(x[0], (pat * (n // 16 + 1))[:n])
        return None

    def imp_sscanf(self):
        """Only the forms thal.c's parameter reader uses: one "%lf" or "%d" conversion."""
        x = self.cpu.x
        s, fmt = self.cpu.cstr(x[0]), self.cpu.cstr(x[1])
        dst = self.cpu.rd(x[31], 8)
        if fmt in (b"%lf", b"%lg", b"%le"):
            m = re.match(rb"\s*([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf(?:inity)?|nan))", s, re.I)
            if not m:
                return 0 if s.strip() else 0xFFFFFFFF
            self.cpu.wr(dst, 8, d2b(float(m.group(1))))
            return 1
        if fmt == b"%d":
            m = re.match(rb"\s*([-+]?\d+)", s)
            if not m:
                return 0 if s.strip() else 0xFFFFFFFF
            self.cpu.wr(dst, 4, int(m.group(1)))
            return 1
        raise NotImplementedError("sscanf(%r, %r)" % (s, fmt))

    def imp_strcat(self):
        x = self.cpu.x
        d = self.cpu.cstr(x[0])
        self._put(x[0] + len(d), self.cpu.cstr(x[1]) + b"\0")
        return x[0]

    def imp_strchr(self):
        x = self.cpu.x
        s = self.cpu.cstr(x[0]) + b"\0"
        i = s.find(bytes([x[1] & 0xFF]))
        return x[0] + i if i >= 0 else 0

    def imp_strcmp(self):
        x = self.cpu.x
        a, b = self.cpu.cstr(x[0]), self.cpu.cstr(x[1])
        return (a > b) - (a < b)

    def imp_strncmp(self):
        x = self.cpu.x
        a, b = self.cpu.cstr(x[0])[: x[2]], self.cpu.cstr(x[1])[: x[2]]
        return (a > b) - (a < b)

    def imp_strcpy(self):
        x = self.cpu.x
        self._put(x[0], self.cpu.cstr(x[1]) + b"\0")
        return x[0]

    def imp_strncpy(self):
        x = self.cpu.x
        s = self.cpu.cstr(x[1])[: x[2]]
        self._put(x[0], s + bytes(x[2] - len(s)))
        return x[0]

    def imp_strlen(self):
        return len(self.cpu.cstr(self.cpu.x[0]))

    def imp_strtod(self):
        x = self.cpu.x
        s = self.cpu.cstr(x[0])
        m = re.match(rb"[ \t\n\v\f\r]*([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf(?:inity)?|nan))", s, re.I)
        if m:
            val, end = float(m.group(1)), m.end()
        else:
            val, end = 0.0, 0
        if x[1]:
            self.cpu.wr(x[1], 8, x[0] + end)
        self.cpu.v[0] = d2b(val)
        return None

    # ---- the further imports primer3_core binds
    def imp___chkstk_darwin(self):
        return None

    def imp__Znwm(self):
        return self.alloc(self.cpu.x[0])

    def imp__ZdlPv(self):
        return None

    def imp__ZNSt3__112__next_primeEm(self):
        n = max(self.cpu.x[0], 2)
        while any(n % d == 0 for d in range(2, int(n ** 0.5) + 1)):
            n += 1
        return n

    def _cxx_runtime(self):
        raise Halt("C++ exception machinery reached")

    imp__Unwind_Resume = imp___cxa_allocate_exception = imp___cxa_throw = imp___gxx_personality_v0 = _cxx_runtime
    imp__ZNSt20bad_array_new_lengthC1Ev = imp__ZNSt20bad_array_new_lengthD1Ev = imp__ZTISt20bad_array_new_length = _cxx_runtime

    def imp_memcpy(self):
        x = self.cpu.x
        self._put(x[0], self._bytes(x[1], x[2]))
        return x[0]

    imp_memmove = imp___memcpy_chk = imp_memcpy

    def imp___strcpy_chk(self):
        return self.imp_strcpy()

    def imp_memchr(self):
        x = self.cpu.x
        i = self._bytes(x[0], x[2]).find(bytes([x[1] & 0xFF]))
        return x[0] + i if i >= 0 else 0

    def imp_abort(self):
        raise Halt("abort(): " + self.out[2].decode(errors="replace"))

    def imp_perror(self):
        self.out[2] += self.cpu.cstr(self.cpu.x[0]) + b": error\n"
        return None

    def _no_file(self):
        return 0xFFFFFFFFFFFFFFFF

    imp_access = imp_stat = imp_open = imp_mmap = _no_file

    def imp_close(self):
        return 0

    imp_munmap = imp_signal = imp_fseek = imp_ftell = imp_close

    def imp_freopen(self):
        return 0

    def _math1(self, fn):
        try:
            r = fn(b2d(self.cpu.v[0]))
        except (ValueError, OverflowError):
            r = math.nan
        self.cpu.v[0] = d2b(r)

    def imp_exp(self):
        a = b2d(self.cpu.v[0])
        try:
            r = math.exp(a)
        except OverflowError:
            r = math.inf
        self.cpu.v[0] = d2b(r)
        return None

    def imp_log10(self):
        a = b2d(self.cpu.v[0])
        self.cpu.v[0] = d2b(math.log10(a) if a > 0 else (-math.inf if a == 0 else math.nan))
        return None

    def imp_pow(self):
        try:
            r = math.pow(b2d(self.cpu.v[0]), b2d(self.cpu.v[1]))
        except (ValueError, OverflowError, ZeroDivisionError):
            r = math.nan
        self.cpu.v[0] = d2b(r)
        return None

    def imp_fputc(self):
        x = self.cpu.x
        self._write(x[1], bytes([x[0] & 0xFF]))
        return x[0] & 0xFF

    def imp_fwrite(self):
        x = self.cpu.x
        self._write(x[3], self._bytes(x[0], x[1] * x[2]))
        return x[2]

    def imp_getline(self):
        x = self.cpu.x
        lineptr, nptr, h = x[0], x[1], x[2]
        if h != 0x10:
            raise IOError("getline on FILE* %#x" % h)
        if self.stdin_pos >= len(self.stdin):
            self.stdin_eof = True
            return 0xFFFFFFFFFFFFFFFF
        e = self.stdin.find(b"\n", self.stdin_pos)
        e = len(self.stdin) if e < 0 else e + 1
        line = self.stdin[self.stdin_pos:e]
        self.stdin_pos = e
        buf, cap = self.cpu.rd(lineptr, 8), self.cpu.rd(nptr, 8)
        if not buf or cap < len(line) + 1:
            cap = max(len(line) + 1, 128)
            buf = self.alloc(cap)
            self.cpu.wr(lineptr, 8, buf)
            self.cpu.wr(nptr, 8, cap)
        self._put(buf, line + b"\0")
        return len(line)

    def imp_getopt_long_only(self):
        return 0xFFFFFFFF          # no options: the reference starts primer3_core without arguments (primer.rs:125-131)

    def imp_qsort(self):
        """qsort(base, nel, width, compar) with the comparison function in emulated code."""
        import functools
        x = self.cpu.x
        base, nel, width, compar = x[0], x[1], x[2], x[3]
        saved = (x[30], x[31])
        items = [self._bytes(base + i * width, width) for i in range(nel)]
        ta, tb = self.alloc(width), self.alloc(width)

        def cmp(a, b):
            self._put(ta, a)
            self._put(tb, b)
            x[0], x[1], x[30] = ta, tb, STOP
            self.cpu.run(compar, STOP)
            return sx(x[0], 32)
        items.sort(key=functools.cmp_to_key(cmp))
        for i, it in enumerate(items):
            self._put(base + i * width, it)
        x[30], x[31] = saved
        return None

    def imp_strncat(self):
        x = self.cpu.x
        d = self.cpu.cstr(x[0])
        self._put(x[0] + len(d), self.cpu.cstr(x[1])[: x[2]] + b"\0")
        return x[0]

    def imp_strtol(self):
        x = self.cpu.x
        s = self.cpu.cstr(x[0])
        m = re.match(rb"[ \t\n\v\f\r]*([-+]?\d+)", s)
        val, end = (int(m.group(1)), m.end()) if m else (0, 0)
        if x[1]:
            self.cpu.wr(x[1], 8, x[0] + end)
        return val & M64

    # ---- entry
    def run_main(self, argv, stdin=b"", limit=None):
        cpu = self.cpu
        self.stdin = stdin
        ptrs = []
        for a in argv:
            p = self.alloc(len(a) + 1)
            self._put(p, a.encode() + b"\0")
            ptrs.append(p)
        av = self.alloc(8 * (len(ptrs) + 1))
        for i, p in enumerate(ptrs):
            cpu.wr(av + 8 * i, 8, p)
        sp = self.base + self.mem_size - 0x10000
        cpu.x[:] = [0] * 32
        cpu.x[0], cpu.x[1], cpu.x[30], cpu.x[31] = len(argv), av, STOP, sp
        code = None
        try:
            cpu.run(self.im.entry, STOP, limit)
            code = sx(cpu.x[0], 32)
        except Exit as e:
            code = e.code
        return self.out[1].decode(), self.out[2].decode(), code


def run_primer3_core(stdin, limit=None):
    """Run `primer3_core` as primer.rs:125-140 does: no arguments, Boulder-IO records on stdin; returns (stdout, stderr, exit
    code, instructions executed)."""
    p = Process(path="/root/reference/od-msspe/bin/primer3_core", mem_size=160 << 20)
    out, err, code = p.run_main(["primer3_core"], stdin, limit)
    return out, err, code, p.cpu.icount


def run_ntthal(args, stdin=b"", file_root=None, limit=None):
    """Run `ntthal <args>`; returns (stdout, stderr, exit code, instructions executed)."""
    p = Process(file_root=file_root)
    out, err, code = p.run_main(["ntthal"] + list(args), stdin, limit)
    return out, err, code, p.cpu.icount


if __name__ == "__main__":
    o, e, c, n = run_ntthal(sys.argv[1:], file_root="/root/reference/od-msspe/primer3_config")
    sys.stdout.write(o)
    sys.stderr.write(e)
    sys.stderr.write("[exit %s, %d instructions]\n" % (c, n))
