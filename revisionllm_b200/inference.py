"""`inference()` - the call the eval drivers make per batch of windows.

Mirrors /root/reference/revisionllm/inference.py:28-75: Vicuna-v1 prompt -> ids with the -200
placeholder -> `model.generate(images=..., query_feats=...)` -> decoded, stripped strings plus the
generate output (`['sequences']`, `['scores']`).  Callers: eval_nlq_negative.py:287,
eval_nlq_retrieval_e2e2.py:353.
"""
from __future__ import annotations

import torch

from .constants import IMAGE_TOKEN_INDEX
from .conversation import SeparatorStyle, conv_templates
from .mm_utils import tokenizer_image_token


def inference(model, image, query_feats, query, tokenizer, visual_memory=None, prefix_memory=None, return_list=False,
              max_new_tokens: int = 1024, output_scores: bool = True, do_sample: bool = False, temperature: float = 0.05,
              seed: int = 0):
    """`do_sample=False` (default) decodes greedily (BASELINE.json north_star); `do_sample=True` is the reference's own rule
    (inference.py:47-48: multinomial at temperature 0.05)."""
    if visual_memory is not None:
        query = query + "<memory>"
    conv = conv_templates["v1"].copy()
    conv.append_message(conv.roles[0], query)
    conv.append_message(conv.roles[1], None)
    prompt = conv.get_prompt()
    input_ids = tokenizer_image_token(prompt, tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt").unsqueeze(0)
    input_ids = input_ids.repeat(image.shape[0], 1)
    stop_str = conv.sep if conv.sep_style != SeparatorStyle.TWO else conv.sep2
    with torch.inference_mode():
        model_output = model.generate(
            input_ids, images=image, query_feats=query_feats, do_sample=do_sample, temperature=temperature, seed=seed, num_beams=1,
            max_new_tokens=max_new_tokens, use_cache=True, visual_memory=visual_memory, prefix_memory=prefix_memory,
            output_scores=output_scores, return_dict_in_generate=True, output_hidden_states=False)
    # the answer = what follows the prompt ids, decoded, with the conversation's stop string and surrounding blanks removed
    generated = model_output["sequences"][:, input_ids.shape[1]:].cpu()
    answers = [text.strip().removesuffix(stop_str).strip() for text in tokenizer.batch_decode(generated, skip_special_tokens=True)]
    single = len(answers) == 1 and not return_list
    return (answers[0] if single else answers), model_output
