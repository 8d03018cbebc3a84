"""Extract the small auxiliary-data fixtures of the reference's rotated-MNIST config.

Run once in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_mnist_fixture.py
Writes tests/golden/mnist_aux.npz with
    pca_ov_init   (400, 8)  PCA object vectors   <- "MNIST data/pca_ov_init3.p"   (MNIST_experiment.py:100-102)
    train_mask    (5400,)   bool                 <- "MNIST data/train_ids_mask3.p" (GPVAE_Casale_model.py:24-38)
    eval_aux      (640, 10) [id, angle, 8 PCA]   <- "MNIST data/eval_data3.p"['aux_data']
    test_aux      (270, 10)                      <- "MNIST data/test_data3.p"['aux_data']
Only auxiliary data (ids, angles, PCA embeddings) is kept -- no images, no reference source.
"""
import os
import pickle

import numpy as np

REF = "/root/reference/MNIST data"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    load = lambda f: pickle.load(open(os.path.join(REF, f), "rb"))
    np.savez_compressed(os.path.join(HERE, "mnist_aux.npz"),
                        pca_ov_init=np.asarray(load("pca_ov_init3.p"), dtype=np.float64),
                        train_mask=np.asarray(load("train_ids_mask3.p"), dtype=bool),
                        eval_aux=np.asarray(load("eval_data3.p")["aux_data"], dtype=np.float64),
                        test_aux=np.asarray(load("test_data3.p")["aux_data"], dtype=np.float64))


if __name__ == "__main__":
    main()
