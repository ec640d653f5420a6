"""Training form of get_outputs (autograd.Functions over the hand-written backward kernels)."""


def get_outputs_train(model, ray_bundle):
    raise NotImplementedError("training path: under construction")
