"""Generate tests/golden/checkpoint_schema.json from the checkpoints shipped with the reference
(/root/reference/paper_pretrained_models/**).  Run in the build container only (the reference
tree does not exist on the GPU box); the JSON is the committed fixture.

    python tests/golden/make_checkpoint_schema.py
"""
import glob
import json
import os

import torch

REF = '/root/reference/paper_pretrained_models'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'checkpoint_schema.json')


def main():
    out = {}
    files = sorted(glob.glob(os.path.join(REF, '**', '*.pt'), recursive=True) +
                   glob.glob(os.path.join(REF, '**', '*.pth.tar'), recursive=True))
    for f in files:
        ck = torch.load(f, map_location='cpu', weights_only=False)
        entry = {'keys': sorted(k for k in ck.keys()),
                 'model': {k: list(v.shape) for k, v in ck['model'].items()},
                 'optimizer_param_groups': [{k: (v if not isinstance(v, (list, tuple)) or k == 'betas' else len(v))
                                             for k, v in g.items()} for g in ck['optimizer']['param_groups']],
                 'optimizer_state_entries': len(ck['optimizer']['state']),
                 'hyper': {k: (v if isinstance(v, (int, float, str, bool, type(None))) else repr(v))
                           for k, v in ck.items() if k not in ('model', 'optimizer')}}
        out[os.path.relpath(f, REF)] = entry
    with open(OUT, 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True, default=str)
    print('wrote', OUT, len(out), 'checkpoints')


if __name__ == '__main__':
    main()
