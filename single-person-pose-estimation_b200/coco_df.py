"""Drop-in for the reference's coco_df.py: COCO person-keypoint annotations -> the merged pandas dataframe gen_TFRecords
iterates (index = image id; columns coco_url, image_path, width, height, ann_id, is_crowd, bbox, num_keypoints, keypoints),
using hgb200.cocoeval.COCO instead of pycocotools."""
from __future__ import annotations

import pandas as pd

from .cocoeval import COCO


def get_meta(coco):
    """coco_df.py:6-21: per image, its file name, size, url and the annotations of every person in it."""
    for img_id in list(coco.imgs.keys()):
        meta = coco.imgs[img_id]
        anns = coco.loadAnns(coco.getAnnIds(imgIds=img_id))
        yield [img_id, meta["file_name"], meta["width"], meta["height"], meta["coco_url"], anns]


def convert_to_df(coco):
    """coco_df.py:23-55 -> (images_df, persons_df), both indexed by image_id."""
    images, persons = [], []
    for img_id, file_name, w, h, url, anns in get_meta(coco):
        images.append({"image_id": int(img_id), "coco_url": url, "image_path": file_name, "width": int(w), "height": int(h)})
        for m in anns:
            persons.append({"ann_id": m["id"], "image_id": m["image_id"], "is_crowd": m["iscrowd"], "bbox": m["bbox"],
                            "num_keypoints": m["num_keypoints"], "keypoints": m["keypoints"]})
    images_df = pd.DataFrame(images)
    images_df.set_index("image_id", inplace=True)
    persons_df = pd.DataFrame(persons)
    persons_df.set_index("image_id", inplace=True)
    return images_df, persons_df


def _filtered(annot_file, min_num_kps):
    images_df, persons_df = convert_to_df(COCO(annot_file))
    df = pd.merge(images_df, persons_df, right_index=True, left_index=True)
    return df[(df["is_crowd"] == 0) & (df["num_keypoints"] >= min_num_kps)]


def gen_trainval_df(config, drop_min_num_kps: bool = False):
    """coco_df.py:57-82: non-crowd people with at least MIN_NUM_KEYPOINTS (or 1) labelled joints, train and valid."""
    min_num_kps = config.MIN_NUM_KEYPOINTS if drop_min_num_kps else 1
    train_df = _filtered(config.TRAIN_ANNOT_FILE, min_num_kps)
    valid_df = _filtered(config.VALID_ANNOT_FILE, min_num_kps)
    print(f"Only examples that are not crowd and num_keypoints >= {min_num_kps} are chosen !")
    print(f"Length of train df: {len(train_df)}")
    print(f"Length of valid df: {len(valid_df)}")
    return train_df, valid_df
