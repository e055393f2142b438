"""Static sweep: decode every word of __text of the executable and list the encodings cpu.py does not handle."""
import sys, collections
sys.path.insert(0, __import__('os').path.dirname(__file__))
from macho import Image
from cpu import CPU, decode, Unknown
path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/od-msspe/bin/ntthal"
im = Image(path)
BASE = im.base
mem = bytearray(0x40000)
im.map_into(mem, BASE)
cpu = CPU(mem, BASE)
addr, sz, _ = im.sections["__text"]
bad = collections.Counter(); ex = {}
for pc in range(addr, addr + sz, 4):
    w = cpu.rd(pc, 4)
    try:
        decode(cpu, w, pc)
    except Unknown as e:
        k = str(e); bad[k] += 1; ex.setdefault(k, []).append((pc, w))
print(sz // 4, "words,", sum(bad.values()), "unknown")
for k, n in bad.most_common():
    print(n, k, " ".join("%x:%08x" % t for t in ex[k][:60]))
