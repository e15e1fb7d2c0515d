class TextualInversionLoaderMixin:
    def maybe_convert_prompt(self, prompt, tokenizer):
        return prompt
