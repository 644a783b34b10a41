"""Input builders shared by gen_golden.py (needs the reference) and the tests (do not)."""
import numpy as np

import cases


def results_fixture():
    """A small synthetic result list (3 images x 80 classes) + the dataset attributes det2json reads."""
    rs = np.random.RandomState(77)
    results = []
    for img in range(3):
        per_class = []
        for c in range(80):
            k = int(rs.randint(0, 3)) if (c + img) % 7 == 0 else 0
            b = cases.random_dets(rs, k, 300, 100) if k else np.zeros((0, 5), np.float32)
            per_class.append(b.astype(np.float32))
        results.append(per_class)

    class DS(object):
        img_ids = [139, 285, 632]
        cat_ids = [i * 2 + 1 for i in range(80)]

        def __len__(self):
            return 3
    return DS(), results
