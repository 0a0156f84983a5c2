"""Development aid: decode the scheduling control bits of `cuobjdump -sass` output (Volta+ 128-bit encoding):
stall cycles, yield, write barrier, read barrier, wait mask.  usage: sass_ctrl.py file.sass [grep-regex]"""
import re, sys
lines = open(sys.argv[1]).read().split('\n')
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
i = 0
while i < len(lines):
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s*/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = hi >> 41
            stall = ctrl & 0xf; yld = (ctrl >> 4) & 1; wr = (ctrl >> 5) & 7; rd = (ctrl >> 8) & 7; wait = (ctrl >> 11) & 0x3f
            s = '%s  %-80s st=%2d y=%d wr=%s rd=%s wait=%s' % (m.group(1), m.group(2).strip()[:80], stall, yld, '-' if wr == 7 else wr, '-' if rd == 7 else rd,
                                                               ''.join(str(b) for b in range(6) if wait >> b & 1) or '-')
            if pat is None or pat.search(s): print(s)
            i += 2
            continue
    i += 1
