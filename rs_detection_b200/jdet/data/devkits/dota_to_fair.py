"""jdet.data.devkits.dota_to_fair, CSV part -- python/jdet/data/devkits/dota_to_fair.py:6-33,102-116.

Reads the merged per-class `after_nms/<Class>.txt` files and writes the FAIR1M-1.5 submission CSV
(`<id>.tif,<Class>,<8 coords %.4f>,<score %.4f>`).  The XML writer (:35-100) is not on the path.
"""
import os


def pick_res(path, images_dir, keep_underline=False):
    """:6-33.  Scenes come from the `*.png` tile names under images_dir (text before the first '__')."""
    res = {}
    for _, _, files in os.walk(images_dir):
        for f in files:
            if f.endswith(".png"):
                res[f.split("__")[0]] = []
    for root, _, files in os.walk(path):
        for f in files:
            cls = f[:-4] if keep_underline else f[:-4].replace("_", " ")
            with open(os.path.join(root, f), "r") as ff:
                for data in ff.read().split("\n"):
                    if len(data) < 5:
                        continue
                    sp = data[:-1].split(" ")  # the reference drops the last character of every line
                    if sp[0] not in res:
                        raise AssertionError(sp[0])
                    res[sp[0]].append({"cls": cls, "p": float(sp[1]), "box": [float(v) for v in sp[2:]]})
    return res


def dota_to_fair1m_1_5(src_path, tar_path, images_dir, name):
    """:102-116"""
    data = pick_res(src_path, images_dir, keep_underline=True)
    os.makedirs(tar_path, exist_ok=True)
    lines = []
    for i in data:
        for obj in data[i]:
            lines.append('{},{},{:.4f},{:.4f},{:.4f},{:.4f},{:.4f},{:.4f},{:.4f},{:.4f},{:.4f}\n'.format(
                str(int(i[1:])) + ".tif", obj["cls"], *obj["box"][:8], obj["p"]))
    with open(os.path.join(tar_path, f"{name}.csv"), "w") as f:
        f.writelines(lines)
