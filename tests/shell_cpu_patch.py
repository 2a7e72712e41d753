"""TEST INFRASTRUCTURE: put the model shell on the CPU inside a spawned worker process (the
permanent twin of the ``shell_on_cpu`` fixture of test_shell_host_logic.py, without monkeypatch:
worker processes are thrown away)."""
import torch


def apply():
    import kernel_standins
    import scvae_b200
    import scvae_b200.engine as E
    import scvae_b200.gmvae_engine as GE
    import scvae_b200.hotloop as H
    import scvae_b200.variational_autoencoder as V
    scvae_b200.kernels = kernel_standins
    for module in (E, GE, H):
        module.K = kernel_standins

    def on_cpu(cls, **forced):
        original = cls.__init__

        def init(self, *args, **kwargs):
            kwargs.update(forced)
            original(self, *args, **kwargs)
            if hasattr(self, "overlap_streams"):
                self.overlap_streams = False
        cls.__init__ = init

    on_cpu(E.VAEEngine, device="cpu", tensor_cores=False)
    on_cpu(GE.GMVAEEngine, device="cpu", tensor_cores=False)
    on_cpu(V.VariationalAutoencoder, device="cpu")
    on_cpu(H.TrainLoop, use_graph=False)
    torch.cuda.synchronize = lambda *a, **k: None
    return kernel_standins
