"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain PyTorch fp32 functional ops, no nn.Module, no reference imports) of the reference
algorithm on the Tango inference hot path:

    unet.py       UNet2DConditionModel.forward          (mustango/diffusers/src/diffusers/models/*)
    schedulers.py DDPMScheduler / DDIMScheduler          (.../schedulers/scheduling_ddpm.py, scheduling_ddim.py)
    vae.py        AutoencoderKL.decode_first_stage       (audioldm/variational_autoencoder/*)
    hifigan.py    Generator.forward + vocoder_infer      (audioldm/hifigan/*)
    pipeline.py   AudioDiffusion.inference + Tango.generate tail   (models.py:210-264, tango.py:43-49)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this
package, and only as the checker / CPU baseline — never on the product path (tango_b200/ has no import of it).

Pinning: the restatement is checked against the REAL reference (imported from /root/reference through
oracle/refshim.py) by oracle/make_golden.py, which also writes the golden vectors under tests/golden/; it is also
checked against the diffusers fork's own known-answer constants (tests/test_oracle_pins.py: sinusoid embedding,
DDPM/DDIM full-loop sums, variance values — mustango/diffusers/tests/...). Parity against the exact pip pins of
diffusers (0.18.2 / 0.20.2, not vendored) is UNPINNED — see SURVEY.md §8c and DESIGN.md.
"""
