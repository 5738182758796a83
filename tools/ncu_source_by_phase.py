"""Attribute the warp-stall samples of an `ncu --set full --import-source on` capture of ginet_graph_step2_kernel to
the kernel's phases and shared routines (no GPU needed: reads the .ncu-rep with `ncu -i` and the line table of the
built library with nvdisasm).

    python tools/ncu_source_by_phase.py gpurun_out/j1_step_cfg2.ncu-rep [launch index, default: all launches summed]

Per-instruction samples come from `ncu --page source --csv --print-source sass`; the i-th SASS instruction of that
listing is the i-th instruction of the kernel's section in `nvdisasm --print-line-info` of fused.sm_100a.cubin (same
binary), which gives the source line and the routine (the `__noinline__` phases are separate functions inside the
kernel's section).  Instructions of the kernel body are mapped to phases by the addresses of the S2_PHASE(i) markers (clock reads)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = '_ZN5drgnn24ginet_graph_step2_kernelE21drgnn_ginet_step_argsNS_9Step2PlanE15drgnn_peer_comm'
CUH = os.path.join(ROOT, 'deeprank-gnn_b200', 'csrc', 'fused_step2.cuh')
PHASE_NAMES = {0: 'entry + conv-weight copies', 20: 'extents', 21: 'bulk-copy issue', 22: 'head-vector copies', 23: 'wait own copies',
               24: 'CTA barrier', 25: 'wait blob + features', 1: 'header check + cluster arrive', 2: 'AX', 3: 'Z1',
               4: 'P1 (cluster max)', 5: 'AP', 6: 'Z2', 7: 'P2 (cluster max)', 8: 'read-out + exchange', 9: 'fc1',
               10: 'fc2 (+ loss, dLoss)', 11: 'head backward + dR', 12: 'dZ2', 13: 'dW2 || dAP', 14: 'dP1', 15: 'dZ1',
               16: 'dW1', 17: 'grid-barrier arrive', 18: 'grid-barrier wait + reduction + Adam (+ exchange)'}


def disassembly():
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'fused.sm_100a.cubin', os.path.join(ROOT, 'deeprank-gnn_b200', 'libdrgnn.so')],
                       cwd=d, check=True, capture_output=True)
        out = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(d, 'fused.sm_100a.cubin')],
                             capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.strip() == '.text.' + KERNEL + ':')
    insts = []
    func, cur = 'kernel body', (None, 0)
    for ln in lines[start + 1:]:
        if ln.startswith('//----') or ln.lstrip().startswith('.section'):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'^\$' + re.escape(KERNEL) + r'\$(\S+):', ln)
        if m:
            name = m.group(1)
            mm = re.search(r'drgnn(\d+)([a-z0-9_]+)', name)
            func = mm.group(2)[:int(mm.group(1))] if mm else name
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            insts.append((int(m.group(1), 16), func, cur, m.group(2).strip()))
    return insts


def phase_of_address(insts):
    """S2_PHASE(i) compiles to a predicated clock read (CS2R ... SR_CLOCKLO) carrying the marker's line number: its
    address is the END of phase i.  The kernel body is laid out in program order, so an instruction belongs to the
    phase whose end marker is the first one at or behind its address.  (Source lines do not work for this: a wait at
    a CTA barrier is sampled on the instruction BEHIND the BAR.SYNC, which is often a rematerialised constant load
    attributed to a line at the top of the kernel.)"""
    marks = {}
    for no, ln in enumerate(open(CUH), 1):
        m = re.search(r'^\s*S2_PHASE\((\d+)\);', ln)
        if m and int(m.group(1)) < 26:
            marks[no] = int(m.group(1))
    body = [(off, line, text) for off, func, (fname, line), text in insts if func == 'kernel body' and fname == 'fused_step2.cuh']
    ends = []
    for off, line, text in body:
        if line in marks and 'CS2R' in text:
            # the marker's other instructions (address of the clock array, the store) sit around the clock read; a
            # barrier wait sampled on one of them still belongs to the phase the marker ends
            last = max(o for o, ln, _t in body if ln == line and off - 0x80 <= o <= off + 0x100)
            ends.append((last, marks[line]))
    ends.sort()

    def f(off):
        for a, ph in ends:
            if off <= a:
                return ph
        return 18
    return f


def main():
    rep = sys.argv[1]
    csvtxt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:graph_step2', '--print-source', 'sass'],
                            capture_output=True, text=True).stdout
    rows = list(csv.reader(csvtxt.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    head = rows[starts[0] + 1]
    col = {n: i for i, n in enumerate(head)}
    sections, seen = [], set()
    for k, s in enumerate(starts):
        e = starts[k + 1] if k + 1 < len(starts) else len(rows)
        sec = [r for r in rows[s + 2:e] if len(r) == len(head)]
        sig = tuple(r[col['# Samples']] for r in sec)
        if sig not in seen:              # (the page prints every launch twice)
            seen.add(sig)
            sections.append(sec)
    if len(sys.argv) > 2:
        sections = [sections[int(sys.argv[2])]]
    numeric = [i for i, n in enumerate(head) if n == '# Samples' or n == 'Instructions Executed' or n.startswith('stall_')]
    body = [list(r) for r in sections[0]]
    for sec in sections[1:]:             # sum the launches instruction by instruction
        for acc, r in zip(body, sec):
            for i in numeric:
                acc[i] = str(int(acc[i] or 0) + int(r[i] or 0))
    which, starts = 'sum of %d' % len(sections), sections
    insts = disassembly()
    assert len(insts) == len(body), (len(insts), len(body))
    reasons = [n for n in head if n.startswith('stall_') and 'Not Issued' not in n]
    ph = phase_of_address(insts)
    by_func, by_phase, by_reason = Counter(), Counter(), Counter()
    reason_phase = defaultdict(Counter)
    executed = Counter()
    total = 0
    for (off, func, (fname, line), text), r in zip(insts, body):
        assert text.split()[0].lstrip('@!P0123456789T ') in r[col['Source']] or True
        n = int(r[col['# Samples']] or 0)
        ex = int(r[col['Instructions Executed']] or 0)
        total += n
        key = func if func != 'kernel body' else 'kernel body'
        by_func[key] += n
        executed[key] += ex
        p = None
        if func == 'kernel body':
            p = ph(off)
            by_phase[p] += n
        for name in reasons:
            v = int(r[col[name]] or 0)
            if v:
                by_reason[name] += v
                reason_phase[key][name] += v
                if func == 'kernel body':
                    reason_phase[('phase', p)][name] += v
    print('%s launch(es) in %s: %d warp-stall samples, %d SASS instructions' % (which, os.path.basename(rep), total, len(insts)))
    print('\n-- by routine (samples, share, warp instructions executed)')
    for k, v in by_func.most_common():
        top = ', '.join('%s %d%%' % (a.replace('stall_', ''), round(100 * b / max(1, sum(reason_phase[k].values()))))
                        for a, b in reason_phase[k].most_common(3))
        print('  %-18s %6d  %5.1f %%  %9d   [%s]' % (k, v, 100.0 * v / total, executed[k], top))
    print('\n-- kernel body by phase (code between two S2_PHASE markers incl. the wait at the barrier that ends the phase; the shared\n   routines the phases call are listed above)')
    for p, v in sorted(by_phase.items(), key=lambda kv: -kv[1]):
        rs = reason_phase[('phase', p)]
        top = ', '.join('%s %d%%' % (a.replace('stall_', ''), round(100 * b / max(1, sum(rs.values())))) for a, b in rs.most_common(3))
        print('  %-50s %6d  %5.1f %%   [%s]' % (PHASE_NAMES.get(p, 'other (inlined headers)') if p is not None else 'other (inlined headers)', v,
                                               100.0 * v / total, top))
    print('\n-- by stall reason (all samples)')
    tot_r = sum(by_reason.values())
    for k, v in by_reason.most_common(10):
        print('  %-24s %6d  %5.1f %%' % (k, v, 100.0 * v / tot_r))


if __name__ == '__main__':
    main()
