"""jdet.data.devkits.data_merge -- python/jdet/data/devkits/data_merge.py.

`prepare_data` writes the per-class `before_nms/<Class>.txt` files (`tile score x1..y4`, four decimals,
:29-48) from the runner's `(result, target)` list; `data_merge` chains it with the device merge
(`result_merge.mergebypoly`).  The FAIR1M-1.5 CSV writer of dota_to_fair.py:102-116 is in
`dota_to_fair.py`.  Zip packaging (:50-101) is shell plumbing and is not mirrored.
"""
import os
import pickle

from .result_merge import mergebypoly

DOTA1_CLASSES = ['plane', 'baseball-diamond', 'bridge', 'ground-track-field', 'small-vehicle', 'large-vehicle', 'ship',
                 'tennis-court', 'basketball-court', 'storage-tank', 'soccer-ball-field', 'roundabout', 'harbor',
                 'swimming-pool', 'helicopter']
FAIR1M_1_5_CLASSES = ['Airplane', 'Ship', 'Vehicle', 'Basketball_Court', 'Tennis_Court', 'Football_Field',
                      'Baseball_Field', 'Intersection', 'Roundabout', 'Bridge']
_CLASSES = {'DOTA': DOTA1_CLASSES, 'DOTA1': DOTA1_CLASSES, 'DOTA1_5': DOTA1_CLASSES + ['container-crane'],
            'DOTA2': DOTA1_CLASSES + ['container-crane', 'airport', 'helipad'], 'FAIR1M_1_5': FAIR1M_1_5_CLASSES}


def get_classes_by_name(name):
    """config/constant.py:207-223 (the oriented datasets of this path)."""
    assert name in _CLASSES
    return _CLASSES[name]


def flip_box(box, target):
    """:14-27 -- undo test-time flips recorded in the target dict.  Arithmetic in float64 (`int - np.float32`
    promotes to float64 under the reference's NumPy 1.x; NumPy 2 would keep float32)."""
    ans = [float(box[i]) for i in range(8)]
    mode = target.get("flip_mode")
    if mode is None:
        return ans
    w, h = target['ori_img_size'][0], target['ori_img_size'][1]
    if 'H' in mode:
        for i in (0, 2, 4, 6):
            ans[i] = w - ans[i]
    if 'V' in mode:
        for i in (1, 3, 5, 7):
            ans[i] = h - ans[i]
    return ans


def prepare_data(result_pkl, save_path, classes):
    """:29-48.  `result_pkl`: path of a pickle or the list itself: [((polys (k,8), scores (k,), labels (k,)),
    {"img_file": ...}), ...] -- exactly what `OrientedHeadTail.get_bboxes(...)` rows give per tile."""
    os.makedirs(save_path, exist_ok=True)
    results = result_pkl
    if isinstance(result_pkl, (str, os.PathLike)):
        with open(result_pkl, 'rb') as f:
            results = pickle.load(f)
    data = {}
    for result, target in results:
        img_name = os.path.splitext(os.path.split(target["img_file"])[-1])[0]
        polys, scores, labels = result
        for bbox, score, label in zip(polys, scores, labels):
            b = flip_box(bbox, target)
            line = '{} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f}\n'.format(img_name, score, *b)
            data.setdefault(classes[int(label)], []).append(line)
    for classname, lines in data.items():
        with open(os.path.join(save_path, classname + '.txt'), 'w') as f:
            f.writelines(lines)


def data_merge(result_pkl, save_path, final_path, dataset_type, nms_threshold_type=0):
    """:50-54"""
    prepare_data(result_pkl, save_path, get_classes_by_name(dataset_type))
    os.makedirs(final_path, exist_ok=True)
    mergebypoly(save_path, final_path, nms_threshold_type)
