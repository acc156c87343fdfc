from ...models._placeholder import placeholder

StableDiffusionPipelineOutput = placeholder("StableDiffusionPipelineOutput")
