-- nn_depthmatch.lua -- drop-in replacements, same class names and constructors as the
-- reference, forward only (the hot path is inference; updateGradInput raises):
--   nn.SpatialMatching(maxh, maxw, full_output)        nnx (opticalflow_model.lua:93)
--   nn.SpatialRadialMatching(hWin)                      radial/radial_opticalflow_network.lua:32-34
--   nn.CascadingAddTable(ratios, trainable, single_beta) CascadingAddTable.lua
--   nn.OutputExtractor(maxh, maxw)                      OutputExtractor.lua
--   nn.DenseMatch(geometry)                             fused SpatialMatching+Minus+SoftMax+extract
--   extractoutput.extractOutput / extractOutputMarginalized   extract_output.cpp:357-366
--   x2yxMulti2(geometry, x)                             opticalflow_model_multiscale.lua:72-81
-- NOT EXECUTED in this repository (no Lua in the build image); see INTEGRATION.md.
require 'torch'
require 'nn'
local ffi = require 'ffi'
local dm = require 'depthmatch_ffi'
local C, ctx = dm.C, dm.ctx

local function nobackward(name)
   return function() error(name .. ': backward is outside the inference hot path') end
end

-- ---------------------------------------------------------------- SpatialMatching
local SpatialMatching, parent = torch.class('nn.SpatialMatching', 'nn.Module')
function SpatialMatching:__init(maxh, maxw, full_output)
   parent.__init(self)
   assert(not full_output, 'full_output=true is never used by the reference')
   self.maxh, self.maxw, self.full_output = maxh or 1, maxw or 1, false
end
function SpatialMatching:updateOutput(input)
   local in1, in2 = input[1], input[2]
   self.output:resize(in1:size(2), in1:size(3), self.maxh, self.maxw)
   dm.check(C.dm_match_volume(ctx, dm.pair(in1, in2), self.maxh, self.maxw, dm.DM_VOLUME_SSD,
                              torch.data(self.output)))
   return self.output
end
SpatialMatching.updateGradInput = nobackward('nn.SpatialMatching')

-- ---------------------------------------------------------------- SpatialRadialMatching
local SpatialRadialMatching, rparent = torch.class('nn.SpatialRadialMatching', 'nn.Module')
function SpatialRadialMatching:__init(hWin)
   rparent.__init(self)
   self.hWin = hWin
end
function SpatialRadialMatching:updateOutput(input)
   local in1, in2 = input[1], input[2]
   self.output:resize(in1:size(2), in1:size(3), self.hWin)
   dm.check(C.dm_match_volume(ctx, dm.pair(in1, in2), self.hWin, 1, dm.DM_VOLUME_SSD,
                              torch.data(self.output)))
   return self.output
end
SpatialRadialMatching.updateGradInput = nobackward('nn.SpatialRadialMatching')

-- ---------------------------------------------------------------- CascadingAddTable (forward)
local CascadingAddTable, cparent = torch.class('nn.CascadingAddTable', 'nn.Module')
function CascadingAddTable:__init(ratios, trainable, single_beta)
   cparent.__init(self)
   self.ratios = ratios
   self.output = {}
   for i = 1, #ratios do self.output[i] = torch.Tensor() end
end
function CascadingAddTable:updateOutput(input)
   if #input ~= #self.ratios then
      error('nn.CascadingAddTable: input and ratios must have the same size')
   end
   local n, rows, kh, kw = #input, input[1]:size(1), input[1]:size(2), input[1]:size(3)
   local stacked = torch.Tensor(n, rows, kh, kw)
   for i = 1, n do
      if input[i]:nDimension() ~= 3 then
         error('nn.CascadingAddTable: input must be a table of 3D-tensors (HxW) x Kh x Kw')
      end
      stacked[i]:copy(input[i])
   end
   local out = torch.Tensor(n, rows, kh, kw)
   local rat = ffi.new('int[?]', n, self.ratios)
   dm.check(C.dm_cascade_add(ctx, torch.data(stacked), rows, kh, kw, rat, n, torch.data(out)))
   for i = 1, n do self.output[i] = out[i] end
   return self.output
end
CascadingAddTable.updateGradInput = nobackward('nn.CascadingAddTable')
function CascadingAddTable:parameters() return {}, {} end
function CascadingAddTable:updateNormalizers() end

-- ---------------------------------------------------------------- OutputExtractor
local OutputExtractor = torch.class('nn.OutputExtractor', 'nn.Module')
function OutputExtractor:__init(maxh, maxw)
   self.maxh, self.maxw = maxh, maxw
   self.output = nil
end
function OutputExtractor:updateOutput(input)
   local inp = input:contiguous()
   local h, w = inp:size(1), inp:size(2)
   local x, y = torch.Tensor(h, w), torch.Tensor(h, w)
   dm.check(C.dm_soft_mean(ctx, torch.data(inp), h * w, self.maxh, self.maxw, torch.data(y), torch.data(x)))
   self.output = {x, y}
   return self.output
end
OutputExtractor.updateGradInput = nobackward('nn.OutputExtractor')

-- ---------------------------------------------------------------- DenseMatch (fused)
-- forward({in1, in2}) returns the table processOutput(geometry, model:forward(input), true, nil)
-- builds (opticalflow_model.lua:201-252), every field the callers read set:
--   index, y, x, confidences, full (2 x hImg x wImg), full_confidences (hImg x wImg)
-- for both extraction methods: 'max' (getOutputConfidences, :153-161: integer flow, confidences 1)
-- and the default 'mean' (getOutputConfidences2, :171-199: soft means, marginal confidence).
local DenseMatch, dparent = torch.class('nn.DenseMatch', 'nn.Module')
function DenseMatch:__init(geometry)
   dparent.__init(self)
   assert(not geometry.multiscale, 'nn.DenseMatch is the single-scale model; use dm_multiscale_extract')
   self.geometry = geometry
end
function DenseMatch:updateOutput(input)
   local g, in1, in2 = self.geometry, input[1], input[2]
   local h, w = in1:size(2), in1:size(3)
   local use_max = g.output_extraction_method == 'max'
   local ret = {index = torch.LongTensor(h, w), pmax = torch.Tensor(h, w)}
   local o = ffi.new('dm_extract_out')
   o.index, o.pmax = torch.data(ret.index), torch.data(ret.pmax)
   local soft, marg
   if use_max then
      ret.full = torch.Tensor(2, g.hImg, g.wImg)
      o.flow_full = torch.data(ret.full)
   else
      soft, marg = torch.Tensor(2, h, w), torch.Tensor(h, w)
      o.soft_yx, o.conf_marginal = torch.data(soft), torch.data(marg)
   end
   dm.check(C.dm_match_extract(ctx, dm.pair(in1, in2), g.maxh, g.maxw, dm.DM_FLAG_TIE_MIDDLE, 0.11,
                               g.hImg, g.wImg, o))
   local yoffset, xoffset = math.ceil(g.maxh / 2), math.ceil(g.maxw / 2)   -- centered2onebased(geometry, 0, 0)
   local hoffset, woffset = math.floor((g.hImg - h) / 2), math.floor((g.wImg - w) / 2)
   if use_max then
      -- the library pasted the integer flow (row - ceil(maxh/2), col - ceil(maxw/2)) into the canvas
      ret.y = ret.full[1]:sub(1 + hoffset, h + hoffset, 1 + woffset, w + woffset):clone()
      ret.x = ret.full[2]:sub(1 + hoffset, h + hoffset, 1 + woffset, w + woffset):clone()
      ret.confidences = torch.Tensor(h, w):fill(1)
   else
      ret.y, ret.x = soft[1] - yoffset, soft[2] - xoffset
      ret.confidences = marg
      -- yx2x(geometry, floor(y + 0.5), floor(x + 0.5)) on the one-based soft means
      local iy, ix = (soft[1] + 0.5):floor(), (soft[2] + 0.5):floor()
      ret.index:copy((iy - 1) * g.maxw + ix)
      ret.full = torch.Tensor(2, g.hImg, g.wImg):zero()
      ret.full:sub(1, 1, 1 + hoffset, h + hoffset, 1 + woffset, w + woffset):copy(ret.y)
      ret.full:sub(2, 2, 1 + hoffset, h + hoffset, 1 + woffset, w + woffset):copy(ret.x)
   end
   ret.full_confidences = torch.Tensor(g.hImg, g.wImg):zero()
   ret.full_confidences:sub(1 + hoffset, h + hoffset, 1 + woffset, w + woffset):copy(ret.confidences)
   self.output = ret
   return ret
end
DenseMatch.updateGradInput = nobackward('nn.DenseMatch')

-- ---------------------------------------------------------------- extractoutput
extractoutput = {}
function extractoutput.extractOutput(input, scores, threshold, ret)
   local inp = input:contiguous()
   dm.check(C.dm_extract_output(ctx, torch.data(inp), inp:size(1), inp:size(2), inp:size(3), threshold,
                                torch.data(ret), torch.data(scores), nil))
end
function extractoutput.extractOutputMarginalized(input, threshold, threshold_acc, ret, retgd)
   local inp = input:contiguous()
   dm.check(C.dm_extract_output_marginalized(ctx, torch.data(inp), inp:size(1), inp:size(2), inp:size(3),
                                             threshold, threshold_acc, torch.data(ret), torch.data(retgd)))
end
package.loaded['extractoutput'] = extractoutput

-- ---------------------------------------------------------------- x2yxMulti2
-- same signature and return order as opticalflow_model_multiscale.lua:72-81; no gcc at run time
function x2yxMulti2(geometry, x, bug_compat)
   local retx = torch.LongTensor():resizeAs(x):zero()
   local rety = torch.LongTensor():resizeAs(x):zero()
   local rat = ffi.new('int[?]', #geometry.ratios, geometry.ratios)
   dm.check(C.dm_x2yx_multi(ctx, torch.data(x), x:size(1), x:size(2), geometry.maxh, geometry.maxw, rat,
                            #geometry.ratios, bug_compat and 1 or 0, torch.data(rety), torch.data(retx)))
   return rety, retx
end

-- ---------------------------------------------------------------- post-processing (no inline.load)
-- opticalflow_model.lua:323-472
function postProcessImage(input, mask, winsize, method)
   local inp, m = input:contiguous(), mask:contiguous()
   local output = torch.Tensor(2, inp:size(2), inp:size(3))
   dm.check(C.dm_post_process_image(ctx, torch.data(inp), torch.data(m), inp:size(2), inp:size(3), winsize,
                                    method == 'max' and 1 or 0, torch.data(output)))
   return output
end

-- depth_estimation_api.lua:76-132 (in place)
function enlargeMask(mask, ix, iy)
   assert(mask:isContiguous())
   dm.check(C.dm_enlarge_mask(ctx, torch.data(mask), mask:size(1), mask:size(2), ix, iy))
   return mask
end

-- test_opticalflow.lua:143-216
function radial(geometry, flow, mh, mw)
   local f = flow:contiguous()
   local h, w = f:size(2), f:size(3)
   local ret, conf = torch.Tensor(h, w), torch.Tensor(h, w)
   dm.check(C.dm_radial_depth(ctx, torch.data(f), h, w, mh or h / 2, mw or w / 2, geometry.wImg / 2,
                              torch.data(ret), torch.data(conf)))
   return ret, conf
end

-- ---------------------------------------------------------------- feature extractor
-- nn.FilterGPU(filter): wraps the nn.Sequential getFilter(geometry) returns
-- (opticalflow_model.lua:45-79) -- the weights stay where load/saveWeights put them, the
-- forward pass of both patches runs on the GPU.  Call :sync() after changing weights.
local FilterGPU, fparent = torch.class('nn.FilterGPU', 'nn.Module')
function FilterGPU:__init(filter)
   fparent.__init(self)
   self.filter = filter
   self.pads = {0, 0, 0, 0}     -- l, r, t, b: the nn.SpatialZeroPadding of getMultiscalePrefilter
   self:sync()
end
function FilterGPU:sync()
   local mods, layers, keep = self.filter.modules, {}, {}
   for i = 1,#mods do
      local m = mods[i]
      if torch.typename(m) == 'nn.Tanh' then
         layers[#layers].tanh_after = 1
      else
         local l = {n_in = m.nInputPlane, n_out = m.nOutputPlane, kh = m.kH, kw = m.kW, n_conn = 0, tanh_after = 0,
                    weight = m.weight:contiguous(), bias = m.bias:contiguous()}
         if m.connTable then
            l.conn = m.connTable:int():contiguous()
            l.n_conn = l.conn:size(1)
         end
         table.insert(layers, l)
      end
   end
   local arr = ffi.new('dm_conv_layer[?]', #layers)
   for i, l in ipairs(layers) do
      local a = arr[i-1]
      a.n_in, a.n_out, a.kh, a.kw, a.n_conn, a.tanh_after = l.n_in, l.n_out, l.kh, l.kw, l.n_conn, l.tanh_after
      a.weight, a.bias = torch.data(l.weight), torch.data(l.bias)
      a.conn = l.conn and torch.data(l.conn) or nil
   end
   local h = ffi.new('dm_filter*[1]')
   dm.check(C.dm_filter_create(ctx, arr, #layers, h))
   self.handle = ffi.gc(h[0], C.dm_filter_destroy)
   self.nOut = layers[#layers].n_out
end
function FilterGPU:updateOutput(input)           -- input: C x H x W or N x C x H x W
   local inp = input:contiguous()
   local n = inp:dim() == 4 and inp:size(1) or 1
   local h, w = inp:size(inp:dim() - 1), inp:size(inp:dim())
   local p = self.pads
   local co, ho, wo = ffi.new('int[1]'), ffi.new('int[1]'), ffi.new('int[1]')
   dm.check(C.dm_filter_output_size(self.handle, h, w, p[1], p[2], p[3], p[4], co, ho, wo))
   if inp:dim() == 4 then self.output:resize(n, co[0], ho[0], wo[0]) else self.output:resize(co[0], ho[0], wo[0]) end
   dm.check(C.dm_filter_forward(ctx, self.handle, torch.data(inp), n, h, w, p[1], p[2], p[3], p[4],
                                torch.data(self.output)))
   return self.output
end
FilterGPU.updateGradInput = nobackward('nn.FilterGPU')
