"""Minimal Mach-O 64 (arm64) image loader for tools/a64emu: segments, LC_SYMTAB, LC_MAIN and LC_DYLD_CHAINED_FIXUPS
(rebases applied in place, binds reported as (address, import name, addend)).  Test infrastructure: it exists so that the
reference's own Primer3 2.6.1 `ntthal` executable (od-msspe/bin/ntthal, spawned by delta_g.rs:90-108) can be executed
instruction by instruction in this x86 image, see emu.py."""
import struct

LC_SEGMENT_64, LC_SYMTAB, LC_MAIN, LC_CHAINED = 0x19, 0x2, 0x80000028, 0x80000034


class Image:
    def __init__(self, path):
        self.b = b = open(path, "rb").read()
        magic, cpu = struct.unpack_from("<Ii", b, 0)
        if magic != 0xFEEDFACF or cpu != 0x0100000C:
            raise ValueError("%s: not a thin Mach-O 64 arm64 executable" % path)
        ncmds, = struct.unpack_from("<I", b, 16)
        off = 32
        self.segments, self.sections, self.syms = [], {}, {}
        self.entry_off, fix = None, None
        for _ in range(ncmds):
            cmd, size = struct.unpack_from("<II", b, off)
            if cmd == LC_SEGMENT_64:
                name = b[off + 8: off + 24].rstrip(b"\0").decode()
                vmaddr, vmsize, fileoff, filesize = struct.unpack_from("<QQQQ", b, off + 24)
                self.segments.append((name, vmaddr, vmsize, fileoff, filesize))
                nsects, = struct.unpack_from("<I", b, off + 64)
                for s in range(nsects):
                    so = off + 72 + s * 80
                    sn = b[so: so + 16].rstrip(b"\0").decode()
                    addr, sz = struct.unpack_from("<QQ", b, so + 32)
                    foff, = struct.unpack_from("<I", b, so + 48)
                    self.sections[sn] = (addr, sz, foff)
            elif cmd == LC_SYMTAB:
                symoff, nsyms, stroff, _ = struct.unpack_from("<IIII", b, off + 8)
                for i in range(nsyms):
                    strx, _t, _s, _d, val = struct.unpack_from("<IBBHQ", b, symoff + 16 * i)
                    nm = b[stroff + strx: b.index(b"\0", stroff + strx)].decode(errors="replace")
                    if val:
                        self.syms[nm] = val
            elif cmd == LC_MAIN:
                self.entry_off, _ = struct.unpack_from("<QQ", b, off + 8)
            elif cmd == LC_CHAINED:
                fix = struct.unpack_from("<II", b, off + 8)
            off += size
        self.base = next(v for n, v, *_ in self.segments if n == "__TEXT")
        self.entry = self.base + self.entry_off
        self.fix = fix

    def map_into(self, mem, mem_base):
        """Copy the segments into `mem` (bytearray whose offset 0 is address mem_base), apply chained rebases, and return
        the binds as a list of (address, symbol, addend)."""
        b = self.b
        for name, vmaddr, vmsize, fileoff, filesize in self.segments:
            if name == "__PAGEZERO" or not filesize:
                continue
            o = vmaddr - mem_base
            mem[o: o + filesize] = b[fileoff: fileoff + filesize]
        binds = []
        if not self.fix:
            return binds
        fo, _ = self.fix
        (_ver, starts_off, imports_off, symbols_off, imports_count, imports_format, _sf) = struct.unpack_from("<7I", b, fo)
        if imports_format != 1:
            raise ValueError("chained import format %d not handled" % imports_format)
        imports = []
        for i in range(imports_count):
            v, = struct.unpack_from("<I", b, fo + imports_off + 4 * i)
            no = v >> 9
            s = fo + symbols_off + no
            imports.append(b[s: b.index(b"\0", s)].decode())
        so = fo + starts_off
        seg_count, = struct.unpack_from("<I", b, so)
        for si in range(seg_count):
            seg_info, = struct.unpack_from("<I", b, so + 4 + 4 * si)
            if not seg_info:
                continue
            p = so + seg_info
            _size, page_size, ptr_format, seg_offset, _maxp, page_count = struct.unpack_from("<IHHQIH", b, p)
            if ptr_format not in (2, 6):
                raise ValueError("chained pointer format %d not handled" % ptr_format)
            for pi in range(page_count):
                start, = struct.unpack_from("<H", b, p + 22 + 2 * pi)
                if start == 0xFFFF:
                    continue
                addr = self.base + seg_offset + pi * page_size + start
                while True:
                    o = addr - mem_base
                    raw, = struct.unpack_from("<Q", mem, o)
                    nxt = (raw >> 51) & 0xFFF
                    if raw >> 63:
                        binds.append((addr, imports[raw & 0xFFFFFF], (raw >> 24) & 0xFF))
                    else:
                        target = raw & 0xFFFFFFFFF
                        high8 = (raw >> 36) & 0xFF
                        if ptr_format == 6:
                            target += self.base
                        struct.pack_into("<Q", mem, o, target | (high8 << 56))
                    if not nxt:
                        break
                    addr += 4 * nxt
        return binds
