"""Extracts the five real `ntthal -a ANY` outputs the reference keeps in its own test (od-msspe/src/delta_g.rs:197-230)
into tests/golden/ntthal_delta_g_rs.json.  Run in the build container (reads /root/reference); the GPU box only
sees the committed JSON.  The Rust source holds the outputs with the tabs already expanded (tab stop 8 in the first
two blocks, 4 in the last three), so the alignment strings are cut at that column."""
import json, re, sys
src = open('/root/reference/od-msspe/src/delta_g.rs').read().split('\n')
lo = next(i for i, l in enumerate(src) if 'pub fn test_parse_ntthal_output' in l)
pairs, blocks, cur = [], [], None
for l in src[lo:lo + 45]:
    t = l.strip()
    m = re.match(r'^([ACGT]{13}),([ACGT]{13})\\n\\$', t)
    if m:
        pairs.append([m.group(1), m.group(2)])
        continue
    body = l[12:] if l.startswith(' ' * 12) else None
    if body is None:
        continue
    body = re.sub(r'(\\n\\|"\.to_string\(\);)$', '', body)
    if body.startswith('Calculated'):
        nums = re.findall(r'(dS|dH|dG|t) = (\S+)', body)
        cur = {'values': {k: v for k, v in nums}, 'lines': []}
        blocks.append(cur)
    elif body[:3] in ('SEQ', 'STR') and cur is not None:
        col = 8 if len(blocks) <= 2 else 4
        cur['lines'].append([body[:3], body[col:]])
assert len(pairs) == 5 and len(blocks) == 5 and all(len(b['lines']) == 4 for b in blocks)
conds = [dict(mv=50, dv=3, dntp=0, dna=250, t=37)] * 2 + [dict(mv=50, dv=3, dntp=0, dna=250, t=25)] * 3
out = [dict(a=p[0], b=p[1], cond=c, **b) for p, c, b in zip(pairs, conds, blocks)]
json.dump(out, open('/root/repo/tests/golden/ntthal_delta_g_rs.json', 'w'), indent=1)
print(json.dumps(out, indent=1)[:1500])
