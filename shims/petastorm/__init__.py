"""Import stub: the parquet reader of the reference's Criteo-1TB loader (recsys/datasets/criteo.py:21) is out of scope
(SURVEY.md section 2, row 7); the npy (Kaggle-format) path is what the synthetic data module feeds."""


def make_batch_reader(*args, **kwargs):
    raise NotImplementedError("petastorm is not available: use the npy (Kaggle-format) datasets")
