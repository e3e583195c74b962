"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's on-disk formats around the merge stage
(SURVEY §8(f) rank 4).  Pure-Python loops, small cases only.  Imported by tests/ and nothing else.

Each function follows the reference lines it cites (paths relative to /root/reference):
  * write_before_nms      python/jdet/data/devkits/data_merge.py:29-48 (+ flip_box :14-27)
  * mergesingle           python/jdet/data/devkits/result_merge.py:206-255 with nmsbynamedict :177-193,
                          poly2origpoly :196-203 and py_cpu_nms_poly_fast :66-127 (-> oracle.py)
  * merge_file            tools/merge_results.py:11-36
  * ensemble_with_class / ensemble_without_class   merge.py:127-176 with nms :14-27 (-> oracle.hbb_nms),
                          poly2obb :73-100 (cv2.minAreaRect, third party: OpenCV) and obb2hbb :103-112
  * fair1m_csv            python/jdet/data/devkits/dota_to_fair.py:6-33,102-116

PARITY UNPINNED for this file: the reference holds no fixtures for these formats and cannot be imported here
(Jittor / Shapely absent); the restatement is checked against hand-written expectations in
tests/test_formats.py only.
"""
import os
import re

import numpy as np

from . import oracle as O

NMS_THRESHOLD_0 = 0.1
NMS_THRESHOLD_1 = {'Roundabout': 0.1, 'Tennis_Court': 0.1, 'Football_Field': 0.1, 'Vehicle': 0.15, 'Ship': 0.2,
                   'Airplane': 0.3, 'Intersection': 0.3, 'Bridge': 0.0001, 'Basketball_Court': 0.1,
                   'Baseball_Field': 0.1}
FAIR1M_1_5_CLASSES = ['Airplane', 'Ship', 'Vehicle', 'Basketball_Court', 'Tennis_Court', 'Football_Field',
                      'Baseball_Field', 'Intersection', 'Roundabout', 'Bridge']


def write_before_nms(results, save_path, classes):
    os.makedirs(save_path, exist_ok=True)
    per_class = {}
    for (polys, scores, labels), target in results:
        stem = os.path.splitext(os.path.split(target["img_file"])[-1])[0]
        for k in range(len(scores)):
            box = [float(v) for v in polys[k]]
            if "flip_mode" in target:
                w, h = target['ori_img_size'][0], target['ori_img_size'][1]
                if 'H' in target["flip_mode"]:
                    box[0::2] = [w - v for v in box[0::2]]
                if 'V' in target["flip_mode"]:
                    box[1::2] = [h - v for v in box[1::2]]
            txt = stem + ' ' + ' '.join('%.4f' % v for v in [float(scores[k])] + box) + '\n'
            per_class.setdefault(classes[int(labels[k])], []).append(txt)
    for name, lines in per_class.items():
        with open(os.path.join(save_path, name + '.txt'), 'w') as f:
            f.write(''.join(lines))


def mergesingle(dstpath, fullname, nms_threshold_type=0, nms="poly", stable_ties=False):
    """nms = "poly": py_cpu_nms_poly_fast (mergebypoly); "rec": py_cpu_nms on 4-coordinate rows (mergebyrec :273-283)."""
    name = os.path.basename(os.path.splitext(fullname)[0])
    by_scene = {}
    with open(fullname) as f:
        for raw in f.readlines():
            parts = raw.strip().split(' ')
            sub = parts[0]
            scene = sub.split('__')[0]
            xy = re.findall(r'\d+', re.findall(r'__\d+___\d+', sub)[0])
            x, y = int(xy[0]), int(xy[1])
            rate = re.findall(r'__([\d+\.]+)__\d+___', sub)[0]
            poly = [float(v) for v in parts[2:]]
            det = O.poly2origpoly(poly, x, y, rate).tolist() + [float(parts[1])]
            by_scene.setdefault(scene, []).append(det)
    thr = NMS_THRESHOLD_0 if nms_threshold_type == 0 else NMS_THRESHOLD_1[name]
    os.makedirs(dstpath, exist_ok=True)
    with open(os.path.join(dstpath, name + '.txt'), 'w') as out:
        for scene, dets in by_scene.items():
            keep = O.py_cpu_nms_poly_fast(np.array(dets), thr, stable_ties) if nms == "poly" else O.py_cpu_nms(np.array(dets), thr)
            for k in keep:
                d = dets[k]
                out.write(scene + ' ' + str(d[-1]) + ' ' + ' '.join(str(v) for v in d[:-1]) + '\n')


def merge_file(src_file, dst_path, nms_thr=0.1, stable_ties=False):
    os.makedirs(dst_path, exist_ok=True)
    by_scene = {}
    with open(src_file) as f:
        for raw in f.readlines():
            parts = raw.strip().split(' ')
            v = [float(t) for t in parts[1:]]
            by_scene.setdefault(parts[0], []).append(v[1:] + v[:1])
    with open(os.path.join(dst_path, os.path.split(src_file)[-1]), 'w') as out:
        for scene, dets in by_scene.items():
            arr = np.array(dets)
            for d in arr[O.py_cpu_nms_poly_fast(arr, nms_thr, stable_ties)].tolist():
                out.write(scene + ' ' + str(d[-1]) + ' ' + ' '.join(str(v) for v in d[:-1]) + '\n')


def _poly2obb_cv(polys):
    import cv2
    out = []
    for quad in polys.reshape(-1, 4, 2).astype(np.float32):
        (x, y), (w, h), ang = cv2.minAreaRect(quad)
        if w >= h:
            ang = -ang
        else:
            w, h, ang = h, w, -90 - ang
        out.append([x, y, w, h, ang / 180 * np.pi])
    return np.array(out).reshape(-1, 5) if out else np.zeros((0, 5))


def _obb2hbb64(obb):
    c, w, h, t = obb[:, :2], obb[:, 2:3], obb[:, 3:4], obb[:, 4:5]
    bias = np.concatenate([np.abs(w / 2 * np.cos(t)) + np.abs(h / 2 * np.sin(t)),
                           np.abs(w / 2 * np.sin(t)) + np.abs(h / 2 * np.cos(t))], axis=-1)
    return np.concatenate([c - bias, c + bias], axis=-1)


def ensemble_with_class(data_list, thresh, stable_ties=False):
    out = []
    for image_id in np.unique(data_list[0][:, 0]):
        dets = np.concatenate([d[d[:, 0] == image_id, :] for d in data_list])
        for ci in range(10):
            t = thresh[FAIR1M_1_5_CLASSES[ci]] if isinstance(thresh, dict) else thresh
            sub = dets[dets[:, -1] == ci + 1]
            prop = np.concatenate([_obb2hbb64(_poly2obb_cv(sub[:, 1:9])), sub[:, 9:10]], axis=1)
            keep = O.hbb_nms(prop, t, stable_ties)
            if len(keep) > 0:
                out.append(sub[np.asarray(keep, dtype=np.int64), :])
    return np.concatenate(out)


def ensemble_without_class(data_list, thresh, stable_ties=False):
    out = []
    for image_id in np.unique(data_list[0][:, 0]):
        dets = np.concatenate([d[d[:, 0] == image_id, :] for d in data_list])
        prop = np.concatenate([_obb2hbb64(_poly2obb_cv(dets[:, 1:9])), dets[:, 9:10]], axis=1)
        keep = O.hbb_nms(prop, thresh, stable_ties)
        if len(keep) > 0:
            out.append(dets[np.asarray(keep, dtype=np.int64), :])
    return np.concatenate(out)


def fair1m_csv(src_path, tar_path, images_dir, name):
    scenes = {}
    for _, _, files in os.walk(images_dir):
        for f in files:
            if f.endswith('.png'):
                scenes[f.split('__')[0]] = []
    for root, _, files in os.walk(src_path):
        for f in files:
            with open(os.path.join(root, f)) as ff:
                for row in ff.read().split('\n'):
                    if len(row) < 5:
                        continue
                    parts = row[:-1].split(' ')
                    scenes[parts[0]].append((f[:-4], float(parts[1]), [float(v) for v in parts[2:]]))
    os.makedirs(tar_path, exist_ok=True)
    with open(os.path.join(tar_path, name + '.csv'), 'w') as out:
        for scene, objs in scenes.items():
            for cls, p, box in objs:
                out.write(','.join([str(int(scene[1:])) + '.tif', cls] + ['%.4f' % v for v in box] + ['%.4f' % p]) + '\n')
