"""Stage the reference's model files for the GPU box (which has no /root/reference): copies the handful of Python files of the
hot path - unmodified - from the reference checkout into baseline/_ref/ (git-ignored, shipped by gpurun), so that
`bench.py --impl reference` and the "reference on this GPU" leg time the REAL GlocalTextPathNavCMT / NavCMT instead of the
oracle port.  Run in the build container:  python tools/make_baseline_ref.py   (also called by __graft_entry__.build())."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('VLN_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')
FILES = [
    'VLN-DUET/map_nav_src/models/vilmodel.py', 'VLN-DUET/map_nav_src/models/ops.py', 'VLN-DUET/map_nav_src/models/transformer.py',
    'VLN-DUET/map_nav_src/models/model.py', 'VLN-DUET/map_nav_src/models/vlnbert_init.py', 'VLN-DUET/map_nav_src/models/graph_utils.py',
    'VLN-HAMT/finetune_src/models/vilmodel_cmt.py', 'VLN-HAMT/finetune_src/models/model_HAMT.py',
    'VLN-HAMT/finetune_src/models/vlnbert_init.py', 'VLN-HAMT/finetune_src/utils/misc.py',
]


def main() -> int:
    if not os.path.isdir(SRC):
        print('no reference checkout at %s: baseline/_ref left as it is' % SRC)
        return 0
    n = 0
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    for d in ('VLN-DUET/map_nav_src/models', 'VLN-HAMT/finetune_src/models', 'VLN-HAMT/finetune_src/utils'):
        init = os.path.join(SRC, d, '__init__.py')
        if os.path.exists(init):
            shutil.copyfile(init, os.path.join(DST, d, '__init__.py'))
    print('staged %d reference files under %s' % (n, DST))
    return 0


if __name__ == '__main__':
    sys.exit(main())
