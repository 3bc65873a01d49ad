"""pipeline.generate (reference pipeline.mojo:13-128) over the C ABI.

`generate` takes the 77x768 context(s) the reference computes at pipeline.mojo:41-53, or - with
`with_clip=True` - the prompt's token ids, which the device CLIP text encoder (clip.mojo:56-109,
SURVEY section 8 row f1) turns into the context; the tokenizer (row f4) stays outside.  Everything
from there on - the denoising loop, CFG, sampler steps, the VAE decode and the final rescale/clamp -
runs on the device."""
from __future__ import annotations

import numpy as np

from .api import Clip, Context, Decoder, Diffusion, Encoder
from .sampler import DDPMSampler, get_time_embedding
from .tokenizer import Tokenizer, prompt_tokens


class Pipeline:
    def __init__(self, ctx: Context, image_size: int = 512, max_images: int = 1, cfg: bool = True, seed: int = 0,
                 weights=None, with_clip: bool = False, clip_vocab: int = 0, clip_layers: int = 0,
                 tokenizer=None, tokenizer_vocab: int = 49408, concat_as_written: bool = True, norm_affine: bool = False):
        if image_size % 32:
            raise ValueError("image_size must be a multiple of 32 (latent side a multiple of 4)")
        self.ctx = ctx
        self.image_size = image_size
        self.side = image_size // 8
        self.cfg = cfg
        self.max_images = max_images
        # norm_affine: every norm of every model owns per-channel weights (real checkpoints; weights.tiny_sd_*_name_map)
        self.norm_affine = norm_affine
        self.diffusion = Diffusion(ctx, self.side, self.side, max_batch=max_images * (2 if cfg else 1), norm_affine=norm_affine)
        self.decoder = Decoder(ctx, self.side, self.side, max_batch=max_images, norm_affine=norm_affine)
        self.clip = Clip(ctx, clip_vocab, clip_layers, norm_affine=norm_affine) if with_clip else None
        self.encoder = None
        # Tokenizer(49408, read_file("tokenizer_clip.bin")), pipeline.mojo:32-37: a path, the file's bytes or a Tokenizer
        self.tokenizer = tokenizer if tokenizer is None or isinstance(tokenizer, Tokenizer) else Tokenizer(tokenizer, tokenizer_vocab)
        self.concat_as_written = concat_as_written
        self._weights, self._seed = weights, seed
        if weights is None:
            self.diffusion.init_random(seed)
            self.decoder.init_random(seed + 1)
            if self.clip:
                self.clip.init_random(seed + 2)
        else:
            self.diffusion.load_weights(weights[0])
            self.decoder.load_weights(weights[1])
            if self.clip:
                self.clip.load_weights(weights[2])

    def close(self):
        for m in (self.diffusion, self.decoder, self.clip, self.encoder):
            if m is not None:
                m.close()

    def encode_tokens(self, tokens) -> np.ndarray:
        """clip.forward(tokens) of pipeline.mojo:45-53: token ids (<= 77, zero-padded) -> (77, 768) context."""
        if self.clip is None:
            raise ValueError("pipeline was created with with_clip=False")
        return self.clip.forward(tokens)

    def _encoder(self):
        """The reference builds Encoder() only when an input image is given (pipeline.mojo:66-67)."""
        if self.encoder is None:
            self.encoder = Encoder(self.ctx, self.side, self.side, max_batch=self.max_images, norm_affine=self.norm_affine)
            if self._weights is None or len(self._weights) < 4:
                self.encoder.init_random(self._seed + 3)
            else:
                self.encoder.load_weights(self._weights[3])
        return self.encoder

    @staticmethod
    def resize_image(img, new_h: int, new_w: int):
        """resize_image, helpers/utils.mojo:372-402: nearest neighbour, source index int(i * old / new).
        Pure indexing of the host-side input image (no arithmetic on pixel values)."""
        img = np.asarray(img)
        _, h, w = img.shape
        if (h, w) == (new_h, new_w):
            return img
        ys = (np.arange(new_h) * (h / new_h)).astype(np.int64)
        xs = (np.arange(new_w) * (w / new_w)).astype(np.int64)
        return img[:, ys][:, :, xs]

    def encode_image(self, input_image, encoder_noise):
        """pipeline.mojo:69-77: resize to image_size, rescale (0,255)->(-1,1) and Encoder.forward, on the device."""
        x = np.asarray(input_image, np.float32)
        squeeze = x.ndim == 3
        if squeeze:
            x, encoder_noise = x[None], np.asarray(encoder_noise)[None]
        x = np.stack([self.resize_image(i, self.image_size, self.image_size) for i in x])
        z = self._encoder().forward(x, encoder_noise, rescale=True)
        return z[0] if squeeze else z

    def schedule(self, inference_steps: int, time_as_written: bool = False, strength=None):
        s = DDPMSampler()
        s.set_inference_timesteps(inference_steps)
        if strength is not None:
            s.set_strength(strength)               # pipeline.mojo:77, sampler.mojo:67-73
        temb = np.stack([get_time_embedding(float(t), time_as_written) for t in s.timesteps])
        return s.timesteps.astype(np.int32), temb.astype(np.float32), s.coefficient_table()

    def generate(self, context, uncond_context=None, cfg_scale: float = 7.5, inference_steps: int = 20,
                 seed_val: int = 0, latents=None, noise=None, decode: bool = True, rescale: bool = True,
                 input_image=None, strength: float = 0.8, encoder_noise=None, start_noise=None):
        """Returns (images (n,3,S,S) in [0,255], latents (n,4,S/8,S/8)).  context (n|1,77,768);
        with CFG pass uncond_context of the same shape.  latents/noise default to seeded N(0,1).
        input_image ((3,h,w) or (n,3,h,w), values 0..255) selects the img2img start of pipeline.mojo:66-79:
        latents = add_noise(Encoder(rescaled image, encoder_noise), timesteps[0]) after set_strength."""
        if not 0.0 <= strength <= 1.0:             # pipeline.mojo:23-29
            raise ValueError("Strength must be between 0 and 1")
        def as_context(c):
            if isinstance(c, str):                      # prompt -> bpe_encode -> token ids (pipeline.mojo:39-53)
                if self.tokenizer is None:
                    raise ValueError("pipeline was created without a tokenizer")
                c = prompt_tokens(c, self.tokenizer, self.concat_as_written)
            c = np.asarray(c)
            if np.issubdtype(c.dtype, np.integer):      # token ids -> device CLIP
                return self.encode_tokens(c)[None]
            c = c.astype(np.float32, copy=False)
            return c[None] if c.ndim == 2 else c

        context = as_context(context)
        n = self.max_images if latents is None else np.asarray(latents).shape[0]
        if input_image is not None:
            n = 1 if np.ndim(input_image) == 3 else np.shape(input_image)[0]
        rng = np.random.default_rng(seed_val)
        if latents is None and input_image is None:
            latents = rng.standard_normal((n, 4, self.side, self.side), dtype=np.float32)
        if noise is None:
            noise = rng.standard_normal((inference_steps, n, 4, self.side, self.side), dtype=np.float32)
        use_cfg = uncond_context is not None
        if use_cfg and not self.cfg:
            raise ValueError("pipeline was created with cfg=False")
        ctx_rows = context
        if use_cfg:
            ctx_rows = np.concatenate([context, as_context(uncond_context)], axis=0)
        ts, temb, coef = self.schedule(inference_steps, strength=None if input_image is None else strength)
        if input_image is not None:
            img = np.asarray(input_image, np.float32)
            img = img[None] if img.ndim == 3 else img
            if encoder_noise is None:
                encoder_noise = rng.standard_normal((n, 4, self.side, self.side), dtype=np.float32)
            if start_noise is None:
                start_noise = rng.standard_normal((n, 4, self.side, self.side), dtype=np.float32)
            if len(ts) == 0:
                raise ValueError("strength leaves no denoising step")
            z = self.encode_image(img, encoder_noise)
            sa, sb = DDPMSampler().add_noise_coefficients(int(ts[0]))
            latents = self.ctx.sampler_add_noise(z, start_noise, sa, sb)
            noise = np.asarray(noise)[inference_steps - len(ts):]
        lat = self.diffusion.generate_latents(latents, ctx_rows, ts, temb, coef, noise, cfg=use_cfg,
                                              cfg_scale=cfg_scale)
        if not decode:
            return None, lat
        return self.decoder.forward(lat, rescale=rescale), lat


def generate(prompt: str, backup_prompt: str = "", strength: float = 0.8, cfg: bool = True, cfg_scale: float = 7.5,
             inference_steps: int = 1, seed_val: int = 0, input_image=None, *, image_size: int = 512, ctx=None,
             tokenizer="tokenizer_clip.bin", weights=None, pipeline: Pipeline | None = None):
    """pipeline.generate (pipeline.mojo:13-128) with the reference's argument list: prompt -> tokenizer -> CLIP
    -> (optional img2img start) -> denoising loop -> Decoder -> (3, image_size, image_size) image in 0..255.
    Weights are random as in the reference unless `weights` = (diffusion, decoder, clip[, encoder]) blobs.
    Pass `pipeline` to reuse the model handles between calls (the reference rebuilds them every time)."""
    if not 0.0 <= strength <= 1.0:                                   # pipeline.mojo:23-29
        raise ValueError("Strength must be between 0 and 1")
    own = pipeline is None
    p = pipeline or Pipeline(ctx or Context(0), image_size=image_size, max_images=1, cfg=cfg, seed=seed_val,
                             weights=weights, with_clip=True, tokenizer=tokenizer)
    try:
        img, _ = p.generate(prompt, backup_prompt if cfg else None, cfg_scale=cfg_scale,
                            inference_steps=inference_steps, seed_val=seed_val, input_image=input_image,
                            strength=strength)
        return img[0]
    finally:
        if own:
            p.close()
