"""One-line summary of bench.py JSON lines:  python tools/bench_summary.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:  # noqa: BLE001
        print(f, 'unreadable:', e)
        continue
    fam = d.get('roofline', {}).get('families', {})
    print(f, 'value %d e2e %d lat %.3f ms' % (d['value'], d['e2e']['value'], d['config'].get('single_stream_ms_per_frame', 0)),
          {k: (round(v['total_ms_per_step'], 3), v['launches_per_step']) for k, v in fam.items()})
