"""Common base of the samplers: same constructor, attributes and quirks as the reference's BaseSampler
(models/samplers/baseSampler.py:7-25), written against that interface."""
import torch


class BaseSampler(torch.nn.Module):
    """Keeps the number of Gibbs steps; the concrete samplers provide the conditional draws.

    Behaviour kept from the reference: `repr()` lists the instance attributes, one "name: value" line each
    (baseSampler.py:12-16), and `get_samples()` hands back the NotImplementedError class instead of raising it
    (baseSampler.py:25), which callers that probe the base class can observe."""

    def __init__(self, n_gibbs_sampling_steps, **module_kwargs):
        torch.nn.Module.__init__(self, **module_kwargs)
        self.n_gibbs_sampling_steps = n_gibbs_sampling_steps

    def __repr__(self):
        return "".join("%s: %s\n" % (name, value) for name, value in vars(self).items())

    def _not_provided(self, what):
        raise NotImplementedError("%s does not implement %s" % (type(self).__name__, what))

    def visible_samples(self, *args, **kwargs):
        self._not_provided("visible_samples")

    def hidden_samples(self, *args, **kwargs):
        self._not_provided("hidden_samples")

    def get_samples(self, *args, **kwargs):
        return NotImplementedError
