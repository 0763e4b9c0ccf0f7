"""Evaluator with the reference's API (disvae/evaluate.py:22-117).  Test losses run on the
CUDA path; the MIG/AAM disentanglement metrics (evaluate.py:119-317) are post-training analysis
and outside this repo's scope: asking for them raises."""
import logging
from collections import defaultdict
from timeit import default_timer

import torch
from tqdm import tqdm

from disvae.utils.modelIO import save_metadata

TEST_LOSSES_FILE = "test_losses.log"
METRICS_FILENAME = "metrics.log"


class Evaluator:
    def __init__(self, model, loss_f, device=torch.device("cpu"), logger=logging.getLogger(__name__),
                 save_dir="results", is_progress_bar=True):
        self.device = device
        self.loss_f = loss_f
        self.model = model.to(self.device)
        self.logger = logger
        self.save_dir = save_dir
        self.is_progress_bar = is_progress_bar
        self.logger.info("Testing Device: {}".format(self.device))

    def __call__(self, data_loader, is_metrics=False, is_losses=True):
        start = default_timer()
        is_still_training = self.model.training
        self.model.eval()
        metric, losses = None, None
        if is_metrics:
            metric = self.compute_metrics(data_loader)
        if is_losses:
            self.logger.info('Computing losses...')
            losses = self.compute_losses(data_loader)
            self.logger.info('Losses: {}'.format(losses))
            save_metadata(losses, self.save_dir, filename=TEST_LOSSES_FILE)
        if is_still_training:
            self.model.train()
        self.logger.info('Finished evaluating after {:.1f} min.'.format((default_timer() - start) / 60))
        return metric, losses

    def compute_losses(self, dataloader):
        """evaluate.py:98-117, including its first-batch-only early return (trap T17)."""
        storer = defaultdict(list)
        for data, _ in tqdm(dataloader, leave=False, disable=not self.is_progress_bar):
            data = data.to(self.device)
            with torch.no_grad():
                try:
                    recon_batch, latent_dist, latent_sample = self.model(data)
                    _ = self.loss_f(data, recon_batch, latent_dist, self.model.training, storer,
                                    latent_sample=latent_sample)
                except ValueError:
                    _ = self.loss_f.call_optimize(data, self.model, None, storer)
            return {k: sum(v) / len(dataloader) for k, v in storer.items()}

    def compute_metrics(self, dataloader):
        raise NotImplementedError("MIG / AAM metrics (reference disvae/evaluate.py:119-317) are outside the "
                                  "B200 hot-path scope of this repository")
