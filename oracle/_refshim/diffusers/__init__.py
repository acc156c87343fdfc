"""One-symbol stand-in for `diffusers`, used ONLY to import the reference's VAE Decoder (which needs nothing from
diffusers but `diffusers.utils.is_torch_version`, easyanimate/vae/ldm/models/omnigen_enc_dec.py:5) in a container
where diffusers is not installed.  Test infrastructure; never on the product path."""
