"""Default option tree for the hot path: the leaves of the reference's options/pix3d/config.yaml that the renderer,
the implicit networks, the losses and the evaluation read (SURVEY.md §5 'config / flags'). The reference's own
utils/options.py keeps working with this package (it only needs attribute access); this module exists so that the
package, its tests and bench.py run where the reference tree is absent."""


class Options(dict):
    """dict with attribute access, nested (same behaviour as the reference's EasyDict for reads/writes)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, Options):
            v = Options(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def default_options(H=224, W=224, device="cuda:0"):
    return Options(
        batch_size=12, image_size=[H, W], H=H, W=W, device=device, seed=0,
        arch=dict(latent_dim_shape=512, latent_dim_rgb=512, enc_network="resnet34", enc_pretrained=False,
                  force_symmetry=True,
                  impl_sdf=dict(beta_init=0.1, proj_latent_dim=64, n_hidden_layers=5, n_channels=64, geometric_init=True,
                                init_sphere_radius=0.5, pos_enc=6, skip_connection=[1, 2], weight_norm=False,
                                eikonal_sample_range=[-1, 1]),
                  impl_rgb=dict(proj_latent_dim=64, n_hidden_layers=3, n_channels=64, pos_enc=6, weight_norm=False)),
        eval=dict(batch_size=1, image_size=[64, 64], vox_res=64, num_points=100000, range=[-0.6, 0.6],
                  f_thresholds=[0.005, 0.01, 0.02, 0.05, 0.1, 0.2]),
        data=dict(k_nearest=5, dataset="pix3d", bgcolor=1),
        render=dict(sampler="uniform", n_samples_uniform=64, rand_sample=512, normal_model="volume"),
        reg=dict(normal_tol=0.2, normal_pow=1, sample_temp=4, n_views=1, mask_mse=0, normal_l1=5),
        loss_weight=dict(eikonal=0.03, render=1, mask=0.5, normal=0.01, nearest_img=1, nearest_mask=0.5,
                         nearest_normal=0.01),
        camera=dict(model="perspective", dist=5, focal=4),
    )
