-- depth_estimation_api_patch.lua -- the replacement of depth_estimation_api.lua:164-170 (the
-- matching path inside nextFrameDepth) by ONE library call.  Everything else in nextFrameDepth
-- stays as it is, including the lines that read the result right after (:171-182):
--     output = poutput.full
--     enlargeMask(mask, math.ceil((geometry.wImg-poutput.y:size(2))/2), math.ceil((geometry.hImg-poutput.y:size(1))/2))
--     mask:cmul(poutput.full_confidences)
-- so the table returned here sets `full`, `y`, `x`, `index`, `confidences` and `full_confidences`
-- exactly as processOutput(geometry, moutput, true, nil) does (opticalflow_model.lua:201-252).
-- NOT EXECUTED in this repository (the build image has no Lua / LuaJIT / Torch7); checked
-- statically by tests/test_abi.py (block structure, fields, C-call argument counts).
--
--   before (depth_estimation_api.lua:164-170):
--      local input   = prepareInput(geometry, last_filtered, filtered)
--      local moutput = model:forward(input)
--      local poutput = processOutput(geometry, moutput, true, nil)
--   after:
--      local poutput = matchAndProcess(geometry, last_filtered, filtered)
require 'nn_depthmatch'

local dense = nil   -- nn.DenseMatch(geometry), built on first use (geometry is a global of the API file)

function matchAndProcess(geometry, last_filtered, filtered)
   if dense == nil then dense = nn.DenseMatch(geometry) end
   local input = prepareInput(geometry, last_filtered, filtered)   -- still narrow()ed views, no copy
   local r = dense:forward(input)                                  -- one dm_match_extract call
   local poutput = {index = r.index, y = r.y, x = r.x, confidences = r.confidences,
                    full = r.full, full_confidences = r.full_confidences}
   return poutput
end

return matchAndProcess
